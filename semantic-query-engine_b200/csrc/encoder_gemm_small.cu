// Encoder linear layers for a HANDFUL of tokens (one query, a few queries): Y = X W^T with the
// operands SWAPPED on the tensor core and the K dimension split across CTAs.
//
// With 16 tokens the 128-row tile of encoder_gemm.cu is 7/8 padding, only N / 64 CTAs stream the
// weights, and each of them re-reads all of X: 10-24 us per layer product, 170 dependent launches,
// 1.6 ms per query.  Here the WEIGHT rows are the M operand (128 output features per CTA, every
// weight byte is fetched exactly once chip-wide) and the tokens are the N operand (n_tok = the
// token rows rounded up to 16, <= 128), so the accumulator is [128 features x n_tok] and nothing is
// padded.  K is split so that ~100+ CTAs stream the weight matrix together (all of a CTA's chunks
// are in flight at once); every CTA writes its fp32 partial tile to a workspace, takes a ticket, and
// the LAST CTA of a feature tile adds the partials in split order (deterministic) and applies the
// same epilogues as the big kernel (bias; Q scale / K / V^T split; + residual -> fp32; gelu -> fp16).
//
// Roofline: HBM (the weights: 2 N K bytes per launch), in practice launch- and latency-bound.
#include "sqe_enc.cuh"

namespace sqe {
namespace enc {

constexpr int kSmallThreads = 192;             // warp 0: TMA, warp 1: MMA, warps 2..5: epilogue
constexpr int kSmallStages = 6;
constexpr int kSmallABytes = 128 * kChunkK * 2;                // 16 KB: 128 weight rows of one K chunk
constexpr int kSmallBBytes = 128 * kChunkK * 2;                // up to 128 token rows
constexpr int kSmallStageBytes = kSmallABytes + kSmallBBytes;
constexpr int kSmallSmemBytes = kSmallStages * kSmallStageBytes + 128 + 16 + 1024;
constexpr int kSmallMaxChunks = kSmallStages;                  // K chunks per CTA (all in flight)

enum { kSEpiSplit = 0, kSEpiResF32 = 1, kSEpiGelu = 2 };

struct SmallArgs {
    int n_tok;                // token rows (multiple of 16, <= 128)
    int m_rows;               // rows that exist (stores are guarded by it)
    int n, k;
    int splits, chunks_per_split;
    const float* bias;
    void* out0;
    int64_t ld0;
    __half* out1;
    int64_t ld1;
    int n_split, q_cols;
    float q_scale;
    const float* residual;
    int64_t ldr;
    const float2* res_stats;  // != null: the residual is LayerNorm(residual) with these row statistics, gamma, beta
    const float* res_gamma;
    const float* res_beta;
    float* partials;          // [n / 128][splits][n_tok][128]
    unsigned* tickets;        // [n / 128], zero between launches
    int epi;
};

__device__ __forceinline__ float gelu_erf_small(float x) {
    const float z = fabsf(x) * 0.70710678118654752440f;
    float t;
    asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(t) : "f"(fmaf(0.3275911f, z, 1.0f)));
    float pl = fmaf(t, 1.061405429f, -1.453152027f);
    pl = fmaf(pl, t, 1.421413741f);
    pl = fmaf(pl, t, -0.284496736f);
    pl = fmaf(pl, t, 0.254829592f);
    float ex;
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(ex) : "f"(z * z * -1.4426950408889634f));
    const float hc = 0.5f * x * (pl * t * ex);
    return x >= 0.0f ? x - hc : hc;
}

__global__ void __launch_bounds__(kSmallThreads, 1)
encoder_gemm_small_kernel(const __grid_constant__ CUtensorMap tmap_w, const __grid_constant__ CUtensorMap tmap_x,
                          const SmallArgs a) {
    extern __shared__ uint8_t smem_raw[];
    const uint32_t raw_addr = ptx::smem_u32(smem_raw);
    const uint32_t base = (raw_addr + 1023u) & ~1023u;
    uint8_t* sm = smem_raw + (base - raw_addr);
    const int warp = __shfl_sync(kFull, static_cast<int>(threadIdx.x >> 5), 0);
    const int lane = threadIdx.x & 31;
    const int tile = blockIdx.x;                               // 128 output features
    const int split = blockIdx.y;
    const int n_chunks = a.chunks_per_split;
    const int kc0 = split * n_chunks;

    const uint32_t bar_full = base + kSmallStages * kSmallStageBytes;      // [kSmallStages]
    const uint32_t bar_done = bar_full + 8 * kSmallStages;                 // accumulator complete
    uint32_t* tmem_ptr_smem = reinterpret_cast<uint32_t*>(sm + kSmallStages * kSmallStageBytes + 128);
    __shared__ unsigned s_last;

    if (warp == 0 && lane == 0) {
        ptx::prefetch_tensormap(&tmap_w);
        ptx::prefetch_tensormap(&tmap_x);
        for (int s = 0; s < kSmallStages; ++s) ptx::mbar_init(bar_full + 8 * s, 1);
        ptx::mbar_init(bar_done, 1);
        ptx::fence_barrier_init();
    }
    if (warp == 1) ptx::tmem_alloc<1>(ptx::smem_u32(tmem_ptr_smem), 128);
    ptx::tc_fence_before();
    __syncthreads();
    ptx::tc_fence_after();
    const uint32_t tmem_base = *tmem_ptr_smem;

    if (warp == 0) {
        if (lane == 0) {
            const uint32_t b_bytes = static_cast<uint32_t>(a.n_tok) * kChunkK * 2;
            for (int c = 0; c < n_chunks; ++c) {               // every chunk has its own stage: no reuse
                const uint32_t sa = base + c * kSmallStageBytes;
                const uint32_t fb = bar_full + 8 * c;
                ptx::mbar_expect_tx(fb, kSmallABytes + b_bytes);
                ptx::tma_load_2d(sa, &tmap_w, (kc0 + c) * kChunkK, tile * 128, fb);
                ptx::tma_load_2d(sa + kSmallABytes, &tmap_x, (kc0 + c) * kChunkK, 0, fb);
            }
        }
    } else if (warp == 1) {
        if (lane == 0) {
            const uint32_t idesc = idesc_f16(128, a.n_tok);
            for (int c = 0; c < n_chunks; ++c) {
                ptx::mbar_wait(bar_full + 8 * c, 0);
                ptx::tc_fence_after();
                const uint32_t sa = base + c * kSmallStageBytes;
                const uint64_t da = make_sw128_desc(sa);
                const uint64_t db = make_sw128_desc(sa + kSmallABytes);
#pragma unroll
                for (int k4 = 0; k4 < kChunkK / kUmmaK; ++k4)
                    ptx::umma_f16<1>(tmem_base, da + 2 * k4, db + 2 * k4, idesc, (c | k4) != 0 ? 1u : 0u);
            }
            ptx::umma_commit(bar_done);
        }
    } else {
        // this thread's output feature and its partial row: [tile][split][token][128 features]
        const int quarter = warp & 3;
        const int f = quarter * 32 + lane;                     // feature within the tile = TMEM lane
        float* part = a.partials + (static_cast<size_t>(tile) * a.splits + split) * a.n_tok * 128 + f;
        ptx::mbar_wait(bar_done, 0);
        ptx::tc_fence_after();
        const uint32_t taddr = tmem_base + (static_cast<uint32_t>(quarter * 32) << 16);
        for (int t0 = 0; t0 < a.n_tok; t0 += 32) {
            uint32_t v[32];
            ptx::tmem_ld_32x32(taddr + t0, v);                 // columns beyond n_tok are never used
            ptx::tmem_wait_ld();
#pragma unroll
            for (int j = 0; j < 32; ++j)
                if (t0 + j < a.n_tok) __stcg(part + static_cast<size_t>(t0 + j) * 128, __uint_as_float(v[j]));
        }
        __threadfence();
    }
    ptx::tc_fence_before();
    __syncthreads();
    if (threadIdx.x == 0) s_last = (atomicAdd(a.tickets + tile, 1u) == static_cast<unsigned>(a.splits - 1)) ? 1u : 0u;
    __syncthreads();
    if (warp == 1) {
        ptx::tc_fence_after();
        ptx::tmem_dealloc<1>(tmem_base, 128);
    }
    if (s_last == 0u) return;
    // ------------------------------------------------------------ the tile's last CTA: reduce + epilogue
    __threadfence();
    if (threadIdx.x == 0) a.tickets[tile] = 0u;               // ready for the next launch
    // 32 threads cover the 128 features of a token row (4 each), the six warps take tokens in turn:
    // per pass every thread has `splits` independent 16-byte loads in flight
    const int fq = (threadIdx.x & 31) * 4;
    const int col = tile * 128 + fq;
    const float4 bias = __ldg(reinterpret_cast<const float4*>(a.bias + col));
    const float* pbase = a.partials + static_cast<size_t>(tile) * a.splits * a.n_tok * 128 + fq;
    const int rows = a.m_rows < a.n_tok ? a.m_rows : a.n_tok;
    for (int t = threadIdx.x >> 5; t < rows; t += kSmallThreads / 32) {
        float4 p[16];
#pragma unroll
        for (int s = 0; s < 16; ++s)
            p[s] = (s < a.splits) ? __ldcg(reinterpret_cast<const float4*>(pbase + (static_cast<size_t>(s) * a.n_tok + t) * 128))
                                  : make_float4(0.f, 0.f, 0.f, 0.f);
        float y[4] = {p[0].x, p[0].y, p[0].z, p[0].w};
#pragma unroll
        for (int s = 1; s < 16; ++s) {                         // split order: deterministic (absent splits add 0)
            y[0] += p[s].x; y[1] += p[s].y; y[2] += p[s].z; y[3] += p[s].w;
        }
        y[0] += bias.x; y[1] += bias.y; y[2] += bias.z; y[3] += bias.w;
        if (a.epi == kSEpiResF32) {
            float4 r = __ldg(reinterpret_cast<const float4*>(a.residual + static_cast<int64_t>(t) * a.ldr + col));
            if (a.res_stats != nullptr) {                      // as the LayerNorm kernel computes it (bit-identical)
                const float2 st = __ldg(a.res_stats + t);
                const float4 g4 = __ldg(reinterpret_cast<const float4*>(a.res_gamma + col));
                const float4 e4 = __ldg(reinterpret_cast<const float4*>(a.res_beta + col));
                r.x = fmaf((r.x - st.x) * st.y, g4.x, e4.x);
                r.y = fmaf((r.y - st.x) * st.y, g4.y, e4.y);
                r.z = fmaf((r.z - st.x) * st.y, g4.z, e4.z);
                r.w = fmaf((r.w - st.x) * st.y, g4.w, e4.w);
            }
            *reinterpret_cast<float4*>(static_cast<float*>(a.out0) + static_cast<int64_t>(t) * a.ld0 + col) =
                make_float4(y[0] + r.x, y[1] + r.y, y[2] + r.z, y[3] + r.w);
        } else if (a.epi == kSEpiSplit && col >= a.n_split) {
#pragma unroll
            for (int e = 0; e < 4; ++e)                        // V^T
                a.out1[static_cast<int64_t>(col + e - a.n_split) * a.ld1 + t] = __float2half_rn(y[e]);
        } else {
            if (a.epi == kSEpiGelu) {
#pragma unroll
                for (int e = 0; e < 4; ++e) y[e] = gelu_erf_small(y[e]);
            } else if (col < a.q_cols) {
#pragma unroll
                for (int e = 0; e < 4; ++e) y[e] *= a.q_scale;
            }
            const __half2 h0 = __floats2half2_rn(y[0], y[1]), h1 = __floats2half2_rn(y[2], y[3]);
            uint2 u;
            u.x = *reinterpret_cast<const uint32_t*>(&h0);
            u.y = *reinterpret_cast<const uint32_t*>(&h1);
            *reinterpret_cast<uint2*>(static_cast<__half*>(a.out0) + static_cast<int64_t>(t) * a.ld0 + col) = u;
        }
    }
}

}  // namespace enc

// workspace: partial tiles + tickets, for the largest product of a layer (n <= 4096, 128 tokens, 16 splits)
int64_t encoder_gemm_small_workspace_bytes() {
    return static_cast<int64_t>(4096 / 128) * 16 * 128 * 128 * 4 + 4096;
}

// rows <= 128 token rows.  Returns -1 for arguments this form does not take (the caller falls back).
int launch_encoder_gemm_small(const void* X, int64_t ldx, const void* W, const float* bias, int64_t m, int n, int k,
                              int epilogue, void* out0, int64_t ld0, void* out1, int64_t ld1, int n_split, int q_cols,
                              float q_scale, const float* residual, int64_t ldr, const float* res_stats,
                              const float* res_gamma, const float* res_beta, void* workspace, int64_t workspace_bytes,
                              cudaStream_t stream) {
    using namespace enc;
    if (m < 1 || m > 128 || n % 128 != 0 || n > 4096 || k % kChunkK != 0 || workspace == nullptr ||
        workspace_bytes < encoder_gemm_small_workspace_bytes())
        return -1;
    SmallArgs a = {};
    a.n_tok = static_cast<int>((m + 15) / 16 * 16);
    a.m_rows = static_cast<int>(m);
    a.n = n;
    a.k = k;
    const int tiles = n / 128, chunks = k / kChunkK;
    int splits = 1;                                            // enough CTAs to stream the weights, <= 6 chunks each
    while ((tiles * splits < 96 || chunks / splits > kSmallMaxChunks) && splits < 16 && chunks % (2 * splits) == 0) splits *= 2;
    if (chunks / splits > kSmallMaxChunks) return -1;
    a.splits = splits;
    a.chunks_per_split = chunks / splits;
    a.bias = bias;
    a.out0 = out0;
    a.ld0 = ld0;
    a.out1 = static_cast<__half*>(out1);
    a.ld1 = ld1;
    a.n_split = (epilogue == kSEpiSplit) ? n_split : n;
    a.q_cols = q_cols;
    a.q_scale = q_scale;
    a.residual = residual;
    a.ldr = ldr;
    a.res_stats = reinterpret_cast<const float2*>(res_stats);
    a.res_gamma = res_gamma;
    a.res_beta = res_beta;
    a.epi = epilogue;
    a.tickets = static_cast<unsigned*>(workspace);
    a.partials = reinterpret_cast<float*>(static_cast<char*>(workspace) + 4096);
    CUtensorMap tw, tx;
    int rc = make_map_2d(&tw, W, static_cast<uint64_t>(k), static_cast<uint64_t>(n), static_cast<uint64_t>(k), 128);
    if (rc != 0) return rc;
    // X: the rows that exist (rows beyond them inside the box read as zeros)
    rc = make_map_2d(&tx, X, static_cast<uint64_t>(k), static_cast<uint64_t>(m), static_cast<uint64_t>(ldx),
                     static_cast<uint32_t>(a.n_tok));
    if (rc != 0) return rc;
    cudaError_t e = cudaFuncSetAttribute(encoder_gemm_small_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                         kSmallSmemBytes);
    if (e != cudaSuccess) { set_error("encoder_gemm_small: smem attribute: %s", cudaGetErrorString(e)); return -2; }
    encoder_gemm_small_kernel<<<dim3(tiles, splits), kSmallThreads, kSmallSmemBytes, stream>>>(tw, tx, a);
    e = cudaGetLastError();
    if (e != cudaSuccess) { set_error("encoder_gemm_small: launch: %s", cudaGetErrorString(e)); return -2; }
    return 0;
}

}  // namespace sqe
