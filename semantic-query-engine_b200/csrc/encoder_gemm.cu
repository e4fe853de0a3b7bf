// Encoder linear layers on tcgen05:  Y[M, N] = X[M, K] . W[N, K]^T  (+ bias, + fused epilogue).
//
// X (activations, fp16 row-major) and W (a torch Linear weight [out, in], fp16 row-major) are both
// K-major, so the operand staging is the scoring kernel's (topk_batched.cu): TMA boxes of
// 64 elements x rows with the 128-byte swizzle, a ring of smem stages, tcgen05.mma issued by one
// thread, two fp32 accumulators in TMEM so that the epilogue of tile i overlaps the main loop of
// tile i + 1.  Persistent CTAs (or CTA pairs) walk the tile list with a fixed stride; tiles that
// share a row block of X are adjacent so X is read from HBM once and W stays in L2.
//
// Two geometries:
// (6 operand stages of 32 KB for pairs, 8 of 24 KB for single CTAs, + 18 KB of epilogue staging)
//   <BN = 256, CG = 2>  a CTA pair computes a 256 x 256 tile (cta_group::2, each CTA stages its 128
//                       rows of X and half of the W rows): the throughput form (ingest batches);
//   <BN = 64,  CG = 1>  128 x 64 tiles: the latency form (a handful of queries: M = 128, so the
//                       number of CTAs that stream W is N / 64 instead of N / 256).
//
// Epilogues (8 warps: two per TMEM lane quarter, each takes half of the tile's columns):
//   kEpiSplit   y = acc + bias;  columns < n_split -> fp16 row-major out0 (columns < q_cols are
//               multiplied by q_scale first: the 1/sqrt(64) of the attention scores, exact in fp16),
//               columns >= n_split -> fp16 TRANSPOSED out1[(col - n_split), row]  (V^T for the
//               attention kernel's second MMA: lanes are consecutive rows, so each column is one
//               coalesced 64-byte store per warp).  n_split = N gives a plain fp16 linear layer.
//   kEpiResF32  y = acc + bias + residual (fp32) -> fp32 out0: the pre-LayerNorm sum.
//   kEpiGelu    y = gelu(acc + bias) (erf form, as BERT) -> fp16 out0.
//
// Roofline: tensor pipe.  Algorithmic FLOPs per launch = 2 M N K.
#include <cstdio>
#include <cstdlib>

#include "sqe_enc.cuh"

namespace sqe {
namespace enc {

constexpr int kGemmThreads = 320;              // TMA, MMA, 8 epilogue warps
constexpr int kABytes = kBM * kChunkK * 2;     // 16 KB: this CTA's 128 rows of one K chunk

template <int BN, int CG>
struct GemmCfg {
    static constexpr int kTileM = kBM * CG;
    static constexpr int kBRows = BN / CG;               // W rows this CTA stages per chunk
    static constexpr int kBBytes = kBRows * kChunkK * 2;
    static constexpr int kStageBytes = kABytes + kBBytes;
    // epilogue staging: per warp HALF a 32 x 32 fp32 strip (16 rows at a time), rows padded to 36
    // floats (16-byte aligned, conflict-free for the row-per-lane 128-bit stores).  The rest of the
    // shared memory is operand stages -- although a sixth stage measurably changes nothing: the main
    // loop waits for data a quarter of the time with five stages AND with six (role timers,
    // profiles/r2b_encoder_gemm_timers.txt), i.e. the L2 -> SM rate binds, not the latency
    static constexpr int kStageRowFloats = 36;
    static constexpr int kStagingBytes = 8 * 16 * kStageRowFloats * 4;       // 18 KB
    static constexpr int kStages = (226 * 1024 - kStagingBytes - 1024) / kStageBytes;     // 6 x 32 KB or 8 x 24 KB
    static constexpr int kTmemCols = 2 * BN;             // two accumulators
    static constexpr int kOffStaging = kStages * kStageBytes;
    static constexpr int kOffBar = kOffStaging + kStagingBytes;
    static constexpr int kOffTmemPtr = kOffBar + (2 * kStages + 4) * 8;
    static constexpr int kSmemBytes = kOffTmemPtr + 16 + 1024;
    static_assert(kSmemBytes <= 227 * 1024, "shared memory budget");
    static_assert(kTmemCols == 128 || kTmemCols == 512, "TMEM allocation must be a power of two");
};

enum { kEpiSplit = 0, kEpiResF32 = 1, kEpiGelu = 2 };

struct GemmArgs {
    int64_t m;                // rows of X that exist (stores are guarded by it)
    int n, k;
    int n_mt, n_nt, n_units;  // tiles along M and N, CTAs (or pairs) in the grid
    const float* bias;        // [n]
    void* out0;
    int64_t ld0;
    __half* out1;             // kEpiSplit: transposed part, [n - n_split, ld1]
    int64_t ld1;
    int n_split, q_cols;
    float q_scale;
    const float* residual;    // kEpiResF32: [m, ldr] fp32
    int64_t ldr;
    // the residual is LayerNorm(residual) when res_stats != null: per row {mean, rstd} (written by the
    // LayerNorm kernel, which then does not have to store its fp32 output at all) + that LayerNorm's
    // gamma / beta; recomputed with the LayerNorm kernel's own operations, so the value is bit-identical
    const float2* res_stats;
    const float* res_gamma;
    const float* res_beta;
    long long* dbg;           // diagnostics: [CTA][8] role timers in cycles (null in production)
};

// gelu(x) = x Phi(x) = 0.5 x (1 + erf(x / sqrt 2)), the erf form BERT uses.  erf by Abramowitz &
// Stegun 7.1.26 (|error| <= 1.5e-7 absolute, i.e. <= 1e-7 |x| on the result -- three orders of
// magnitude below the fp16 rounding of the output): with z = |x| / sqrt 2, t = 1 / (1 + p z),
// 1 - erf(z) = t (a1 + t (a2 + t (a3 + t (a4 + t a5)))) exp(-z^2) =: c, and
// gelu(x) = x - x c / 2 for x >= 0, x c / 2 for x < 0.  Two MUFU (rcp, ex2) + 12 FMA-pipe
// instructions per element; erff() costs ~35, which made the FFN epilogue longer than its main loop.
__device__ __forceinline__ float gelu_erf(float x) {
    const float z = fabsf(x) * 0.70710678118654752440f;
    float t;
    asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(t) : "f"(fmaf(0.3275911f, z, 1.0f)));
    float pl = fmaf(t, 1.061405429f, -1.453152027f);
    pl = fmaf(pl, t, 1.421413741f);
    pl = fmaf(pl, t, -0.284496736f);
    pl = fmaf(pl, t, 0.254829592f);
    float ex;
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(ex) : "f"(z * z * -1.4426950408889634f));
    const float hc = 0.5f * x * (pl * t * ex);             // x c / 2
    return x >= 0.0f ? x - hc : hc;
}

__device__ __forceinline__ uint32_t pack_half2(float a, float b) {
    const __half2 h = __floats2half2_rn(a, b);
    return *reinterpret_cast<const uint32_t*>(&h);
}

// CL = CTAs per cluster.  CL = CG: every CTA (pair) is on its own.  CL = 4 (CG = 2): two pairs work on
// neighbouring n-tiles of the same row block and SHARE the X tile: each CTA fetches 64 of the 128 X
// rows it needs and multicasts them to itself and to the CTA of the other pair that needs the same rows
// (24 KB instead of 32 KB out of L2 per CTA and K chunk: the 256 x 256 pair tile is bound by the
// L2 -> SM bandwidth, not by the tensor pipe).  A stage is then written by two CTAs, so its "empty"
// barrier waits for the MMAs of BOTH pairs (their commits are multicast to all four CTAs).
template <int BN, int CG, int EPI, int CL>
__global__ void __launch_bounds__(kGemmThreads, 1)
encoder_gemm_kernel(const __grid_constant__ CUtensorMap tmap_x, const __grid_constant__ CUtensorMap tmap_w,
                    const __grid_constant__ CUtensorMap tmap_xh, const GemmArgs a) {
    static_assert(CL == CG || ((CL == 4 || CL == 8) && CG == 2), "cluster forms");
    constexpr int kPairs = (CL > CG) ? CL / 2 : 1;                   // pairs that share an X tile
    using C = GemmCfg<BN, CG>;
    constexpr int kStages = C::kStages;
    extern __shared__ uint8_t smem_raw[];
    const uint32_t raw_addr = ptx::smem_u32(smem_raw);
    const uint32_t base = (raw_addr + 1023u) & ~1023u;         // SWIZZLE_128B atoms are 1024-B aligned
    uint8_t* sm = smem_raw + (base - raw_addr);

    const int warp = __shfl_sync(kFull, static_cast<int>(threadIdx.x >> 5), 0);
    const int lane = threadIdx.x & 31;
    const uint32_t crank = (CG == 2) ? ptx::cluster_ctarank() : 0u;  // rank in the cluster
    const uint32_t rank = crank & 1u;                                // rank in the pair; 0 = leader (issues the MMAs)
    const uint32_t pair = crank >> 1;                                // CL > 2: which of the pairs
    const uint32_t leader = crank & ~1u;                             // cluster rank of this pair's leader
    const int unit = blockIdx.x / CL;
    // CL > 2: a unit's tile = (row block, kPairs neighbouring n-tiles); this pair takes n-tile kPairs j + pair
    const int nt_per_unit = kPairs;
    const int n_tiles = a.n_mt * (a.n_nt / nt_per_unit);
    const int n_chunks = a.k / kChunkK;
    auto m_tile = [&](int t) { return t / (a.n_nt / nt_per_unit); };
    auto n_tile = [&](int t) { return (t % (a.n_nt / nt_per_unit)) * nt_per_unit + ((CL > CG) ? static_cast<int>(pair) : 0); };

    const uint32_t bar_full = base + C::kOffBar;               // [kStages]
    const uint32_t bar_empty = bar_full + 8 * kStages;         // [kStages]
    const uint32_t bar_tfull = bar_empty + 8 * kStages;        // [2]
    const uint32_t bar_tempty = bar_tfull + 16;                // [2]
    uint32_t* tmem_ptr_smem = reinterpret_cast<uint32_t*>(sm + C::kOffTmemPtr);

    if (warp == 0 && lane == 0) {
        ptx::prefetch_tensormap(&tmap_x);
        ptx::prefetch_tensormap(&tmap_w);
        if constexpr (CL > CG) ptx::prefetch_tensormap(&tmap_xh);
        for (int s = 0; s < kStages; ++s) {
            ptx::mbar_init(bar_full + 8 * s, 1);               // the leader's expect_tx arrival
            ptx::mbar_init(bar_empty + 8 * s, CL / CG);        // one tcgen05.commit per pair that writes this stage
        }
        for (int acc = 0; acc < 2; ++acc) {
            ptx::mbar_init(bar_tfull + 8 * acc, 1);            // one tcgen05.commit
            ptx::mbar_init(bar_tempty + 8 * acc, 8 * CG);      // one arrival per epilogue warp (of both CTAs)
        }
        ptx::fence_barrier_init();
    }
    if (warp == 1) ptx::tmem_alloc<CG>(ptx::smem_u32(tmem_ptr_smem), C::kTmemCols);
    ptx::tc_fence_before();
    if constexpr (CG == 2) ptx::cluster_sync_all(); else __syncthreads();
    ptx::tc_fence_after();
    const uint32_t tmem_base = *tmem_ptr_smem;

    if (warp == 0) {
        // ------------------------------------------------------------ TMA producer
        if (lane == 0) {
            const uint64_t pol_w = ptx::policy_evict_last();   // the weights are re-read by every row block
            int stage = 0;
            uint32_t phase = 0;
            long long t_wait = 0;
            const long long t_begin = a.dbg ? clock64() : 0;
            for (int t = unit; t < n_tiles; t += a.n_units) {
                const int x_row = m_tile(t) * C::kTileM + static_cast<int>(rank) * kBM;
                const int w_row = n_tile(t) * BN + static_cast<int>(rank) * C::kBRows;
                for (int kc = 0; kc < n_chunks; ++kc) {
                    const long long w0 = a.dbg ? clock64() : 0;
                    ptx::mbar_wait(bar_empty + 8 * stage, phase ^ 1u);
                    if (a.dbg) t_wait += clock64() - w0;
                    const uint32_t sa = base + stage * C::kStageBytes;
                    if constexpr (CG == 1) {
                        const uint32_t fb = bar_full + 8 * stage;
                        ptx::mbar_expect_tx(fb, C::kStageBytes);
                        ptx::tma_load_2d(sa, &tmap_x, kc * kChunkK, x_row, fb);
                        ptx::tma_load_2d_hint(sa + kABytes, &tmap_w, kc * kChunkK, w_row, fb, pol_w);
                    } else {
                        // both CTAs' bytes are counted on the LEADER's barrier
                        if (rank == 0) ptx::mbar_expect_tx(bar_full + 8 * stage, 2 * C::kStageBytes);
                        const uint32_t fb = ptx::mapa(bar_full + 8 * stage, leader);
                        if constexpr (CL > CG) {
                            // my share of the 128 X rows, to me and to the same-rank CTAs of the other pairs
                            constexpr uint16_t kEven = (CL == 8) ? 0x55 : 0x5;
                            const uint16_t mask = static_cast<uint16_t>(kEven << rank);
                            ptx::tma_load_2d_cg2_mc(sa + pair * (kABytes / kPairs), &tmap_xh, kc * kChunkK,
                                                    x_row + static_cast<int>(pair) * (kBM / kPairs), fb, mask);
                        } else {
                            ptx::tma_load_2d_cg2(sa, &tmap_x, kc * kChunkK, x_row, fb);
                        }
                        ptx::tma_load_2d_cg2(sa + kABytes, &tmap_w, kc * kChunkK, w_row, fb);
                    }
                    if (++stage == kStages) { stage = 0; phase ^= 1u; }
                }
            }
            if (a.dbg) {
                a.dbg[blockIdx.x * 8 + 0] = clock64() - t_begin;          // producer: total, waiting for free stages
                a.dbg[blockIdx.x * 8 + 1] = t_wait;
            }
        }
    } else if (warp == 1) {
        // -------------------------------------------------------------- MMA issuer
        if (lane == 0 && rank == 0) {
            constexpr uint32_t idesc = idesc_f16(C::kTileM, BN);
            int stage = 0;
            uint32_t phase = 0;
            int i = 0;
            long long t_wfull = 0, t_wtempty = 0;
            const long long t_begin = a.dbg ? clock64() : 0;
            for (int t = unit; t < n_tiles; t += a.n_units, ++i) {
                const int acc = i & 1;
                const long long w0 = a.dbg ? clock64() : 0;
                ptx::mbar_wait(bar_tempty + 8 * acc, ((i >> 1) & 1) ^ 1u);   // the epilogue drained it
                if (a.dbg) t_wtempty += clock64() - w0;
                ptx::tc_fence_after();
                const uint32_t tmem_d = tmem_base + acc * BN;
                for (int kc = 0; kc < n_chunks; ++kc) {
                    const long long w1 = a.dbg ? clock64() : 0;
                    ptx::mbar_wait(bar_full + 8 * stage, phase);        // TMA bytes have landed
                    if (a.dbg) t_wfull += clock64() - w1;
                    ptx::tc_fence_after();
                    const uint32_t sa = base + stage * C::kStageBytes;
                    const uint64_t da = make_sw128_desc(sa);
                    const uint64_t db = make_sw128_desc(sa + kABytes);
#pragma unroll
                    for (int k4 = 0; k4 < kChunkK / kUmmaK; ++k4) {
                        // advance 16 elements = 32 bytes inside the swizzle row: +2 in 16-B units
                        ptx::umma_f16<CG>(tmem_d, da + 2 * k4, db + 2 * k4, idesc, (kc | k4) != 0 ? 1u : 0u);
                    }
                    if constexpr (CG == 1) ptx::umma_commit(bar_empty + 8 * stage);
                    else ptx::umma_commit_cg2(bar_empty + 8 * stage, static_cast<uint16_t>((1u << CL) - 1u));     // every CTA that writes it
                    if (++stage == kStages) { stage = 0; phase ^= 1u; }
                }
                if constexpr (CG == 1) ptx::umma_commit(bar_tfull + 8 * acc);
                else ptx::umma_commit_cg2(bar_tfull + 8 * acc, static_cast<uint16_t>(0x3u << leader));
            }
            if (a.dbg) {
                a.dbg[blockIdx.x * 8 + 2] = clock64() - t_begin;          // MMA issuer: total, waiting for data, for the epilogue
                a.dbg[blockIdx.x * 8 + 3] = t_wfull;
                a.dbg[blockIdx.x * 8 + 4] = t_wtempty;
                a.dbg[blockIdx.x * 8 + 5] = i;                             // tiles
                unsigned long long gt;
                asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(gt));
                a.dbg[blockIdx.x * 8 + 6] = static_cast<long long>(gt);
            }
        }
    } else {
        // ---------------------------------------------------------------- epilogue
        // TMEM gives every thread one ROW of the tile (32 consecutive columns per load).  Stored
        // as such, a warp-wide 128-bit access touches 32 different lines (16 useful bytes each) and
        // the LSU, not the tensor pipe, bounds the narrow-K layers.  So each warp transposes its
        // 32 x 32 strip through shared memory and reads / writes global memory row-contiguously:
        // eight lanes cover one 128-byte row segment, a warp instruction covers four rows.
        const int quarter = warp & 3;                          // TMEM lanes 32 q .. 32 q + 31
        const int half = (warp - 2) >> 2;                      // which half of the tile's columns
        constexpr int kStrips = BN / 64;                       // strips of 32 columns per warp
        constexpr int RS = C::kStageRowFloats;
        float* stg = reinterpret_cast<float*>(sm + C::kOffStaging) + (warp - 2) * 16 * RS;
        const int sub_row = lane >> 3;                         // row of a 4-row group in the coalesced phase
        const int c4 = (lane & 7) * 4;                         // first of this lane's 4 columns there
        int i = 0;
        for (int t = unit; t < n_tiles; t += a.n_units, ++i) {
            const int acc = i & 1;
            const int64_t row0 = static_cast<int64_t>(m_tile(t)) * C::kTileM + rank * kBM + quarter * 32;
            const int col_t = n_tile(t) * BN + half * (BN / 2);
            [[maybe_unused]] float2 rst[8];                   // {mean, rstd} of this lane's 8 rows (every strip of the tile)
            if constexpr (EPI == kEpiResF32) {
                if (a.res_stats != nullptr) {
#pragma unroll
                    for (int g = 0; g < 8; ++g) {
                        const int64_t row = row0 + 4 * g + sub_row;
                        rst[g] = (row < a.m) ? __ldg(a.res_stats + row) : make_float2(0.f, 0.f);
                    }
                }
            }
            ptx::mbar_wait(bar_tfull + 8 * acc, (i >> 1) & 1);
            ptx::tc_fence_after();
            const uint32_t taddr = tmem_base + (static_cast<uint32_t>(quarter * 32) << 16) + acc * BN + half * (BN / 2);
#pragma unroll 1
            for (int s = 0; s < kStrips; ++s) {
                const int col0 = col_t + 32 * s;
                uint32_t v[32];
                ptx::tmem_ld_32x32(taddr + 32 * s, v);
                if (EPI == kEpiSplit && col0 >= a.n_split) {
                    // V^T: lanes are consecutive rows -> every column is one 64-byte run already
                    const int64_t row = row0 + lane;
                    const float4* bp = reinterpret_cast<const float4*>(a.bias + col0);
                    ptx::tmem_wait_ld();
                    if (row < a.m) {
                        __half* op = a.out1 + static_cast<int64_t>(col0 - a.n_split) * a.ld1 + row;
#pragma unroll
                        for (int j = 0; j < 8; ++j) {
                            const float4 b4 = __ldg(bp + j);
                            op[static_cast<int64_t>(4 * j + 0) * a.ld1] = __float2half_rn(__uint_as_float(v[4 * j + 0]) + b4.x);
                            op[static_cast<int64_t>(4 * j + 1) * a.ld1] = __float2half_rn(__uint_as_float(v[4 * j + 1]) + b4.y);
                            op[static_cast<int64_t>(4 * j + 2) * a.ld1] = __float2half_rn(__uint_as_float(v[4 * j + 2]) + b4.z);
                            op[static_cast<int64_t>(4 * j + 3) * a.ld1] = __float2half_rn(__uint_as_float(v[4 * j + 3]) + b4.w);
                        }
                    }
                    continue;
                }
                // this lane's bias (its 4 columns of every row) and, for the residual form, the
                // residual values of its 8 rows: issued before the TMEM data is awaited
                const float4 b4 = __ldg(reinterpret_cast<const float4*>(a.bias + col0 + c4));
                [[maybe_unused]] float4 r4[8];
                if constexpr (EPI == kEpiResF32) {
#pragma unroll
                    for (int g = 0; g < 8; ++g) {
                        const int64_t row = row0 + 4 * g + sub_row;
                        r4[g] = (row < a.m) ? __ldg(reinterpret_cast<const float4*>(a.residual + row * a.ldr + col0 + c4))
                                            : make_float4(0.f, 0.f, 0.f, 0.f);
                    }
                    if (a.res_stats != nullptr) {              // residual = LayerNorm(residual), as the LN kernel computes it
                        const float4 g4 = __ldg(reinterpret_cast<const float4*>(a.res_gamma + col0 + c4));
                        const float4 e4 = __ldg(reinterpret_cast<const float4*>(a.res_beta + col0 + c4));
#pragma unroll
                        for (int g = 0; g < 8; ++g) {
                            r4[g].x = fmaf((r4[g].x - rst[g].x) * rst[g].y, g4.x, e4.x);
                            r4[g].y = fmaf((r4[g].y - rst[g].x) * rst[g].y, g4.y, e4.y);
                            r4[g].z = fmaf((r4[g].z - rst[g].x) * rst[g].y, g4.z, e4.z);
                            r4[g].w = fmaf((r4[g].w - rst[g].x) * rst[g].y, g4.w, e4.w);
                        }
                    }
                }
                ptx::tmem_wait_ld();
                const float qs = (EPI == kEpiSplit && col0 < a.q_cols) ? a.q_scale : 1.0f;
#pragma unroll
                for (int hp = 0; hp < 2; ++hp) {               // rows 0..15, then 16..31 of the strip
                    if ((lane >> 4) == hp) {
                        float* srow = stg + (lane & 15) * RS;
#pragma unroll
                        for (int j = 0; j < 8; ++j)
                            *reinterpret_cast<uint4*>(srow + 4 * j) = make_uint4(v[4 * j], v[4 * j + 1], v[4 * j + 2], v[4 * j + 3]);
                    }
                    __syncwarp();
#pragma unroll
                    for (int gg = 0; gg < 4; ++gg) {
                        const int g = 4 * hp + gg;
                        const int64_t row = row0 + 4 * g + sub_row;
                        float4 y = *reinterpret_cast<const float4*>(stg + (4 * gg + sub_row) * RS + c4);
                        y.x += b4.x; y.y += b4.y; y.z += b4.z; y.w += b4.w;
                        if constexpr (EPI == kEpiResF32) {
                            y.x += r4[g].x; y.y += r4[g].y; y.z += r4[g].z; y.w += r4[g].w;
                            if (row < a.m)
                                *reinterpret_cast<float4*>(static_cast<float*>(a.out0) + row * a.ld0 + col0 + c4) = y;
                        } else {
                            if constexpr (EPI == kEpiGelu) {
                                y.x = gelu_erf(y.x); y.y = gelu_erf(y.y); y.z = gelu_erf(y.z); y.w = gelu_erf(y.w);
                            } else {
                                y.x *= qs; y.y *= qs; y.z *= qs; y.w *= qs;
                            }
                            if (row < a.m)
                                *reinterpret_cast<uint2*>(static_cast<__half*>(a.out0) + row * a.ld0 + col0 + c4) =
                                    make_uint2(pack_half2(y.x, y.y), pack_half2(y.z, y.w));
                        }
                    }
                    __syncwarp();                              // the buffer is rewritten by the next half / strip
                }
            }
            ptx::tc_fence_before();
            __syncwarp();
            if (lane == 0) {
                if constexpr (CG == 1) ptx::mbar_arrive(bar_tempty + 8 * acc);
                else ptx::mbar_arrive_cluster(bar_tempty + 8 * acc, leader);   // the pair leader's barrier
            }
        }
    }

    // Teardown.  In pair mode neither CTA may exit (or free TMEM) while the other still reads
    // its shared memory / signals its barriers.
    ptx::tc_fence_before();
    if constexpr (CG == 2) ptx::cluster_sync_all(); else __syncthreads();
    if (warp == 1) {
        ptx::tc_fence_after();
        ptx::tmem_dealloc<CG>(tmem_base, C::kTmemCols);
    }
}

template <int BN, int CG, int EPI, int CL>
static int launch_gemm_t(const void* X, const void* W, GemmArgs a, int64_t ldx, int sm_count, cudaStream_t stream) {
    using C = GemmCfg<BN, CG>;
    auto kernel = encoder_gemm_kernel<BN, CG, EPI, CL>;
    a.n_mt = static_cast<int>((a.m + C::kTileM - 1) / C::kTileM);
    a.n_nt = a.n / BN;
    const int tiles = a.n_mt * a.n_nt / (CL / CG);             // a 4-CTA cluster takes two n-tiles at a time
    cudaError_t e = cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, C::kSmemBytes);
    if (e != cudaSuccess) { set_error("encoder_gemm: smem attribute: %s", cudaGetErrorString(e)); return -2; }
    cudaLaunchConfig_t cfg = {};
    cfg.blockDim = dim3(kGemmThreads);
    cfg.dynamicSmemBytes = C::kSmemBytes;
    cfg.stream = stream;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = CL;
    attr[0].val.clusterDim.y = 1;
    attr[0].val.clusterDim.z = 1;
    cfg.attrs = attr;
    cfg.numAttrs = 1;
    int units = sm_count / CL;
    if constexpr (CL > CG) {
        // larger clusters do not tile every GPC: ask how many are co-resident (cached per device)
        static int active[64] = {0};
        int dev = 0;
        cudaGetDevice(&dev);
        if (dev >= 0 && dev < 64) {
            if (active[dev] == 0) {
                cfg.gridDim = dim3(static_cast<unsigned>(sm_count / CL * CL));
                int n = 0;
                if (cudaOccupancyMaxActiveClusters(&n, kernel, &cfg) == cudaSuccess && n > 0) active[dev] = n;
                else { active[dev] = -1; (void)cudaGetLastError(); }
                if (getenv("SQE_DEBUG_CLUSTERS")) fprintf(stderr, "encoder_gemm: clusters of %d co-resident: %d\n", CL, active[dev]);
            }
            if (active[dev] > 0) units = active[dev];
        }
    }
    if (units > tiles) units = tiles;
    a.n_units = units;
    CUtensorMap tx, tw, txh;
    int rc = make_map_2d(&tx, X, static_cast<uint64_t>(a.k), static_cast<uint64_t>(a.m), static_cast<uint64_t>(ldx), kBM);
    if (rc != 0) return rc;
    rc = make_map_2d(&tw, W, static_cast<uint64_t>(a.k), static_cast<uint64_t>(a.n), static_cast<uint64_t>(a.k), C::kBRows);
    if (rc != 0) return rc;
    rc = make_map_2d(&txh, X, static_cast<uint64_t>(a.k), static_cast<uint64_t>(a.m), static_cast<uint64_t>(ldx),
                     kBM / (CL > CG ? CL / 2 : 1));
    if (rc != 0) return rc;
    cfg.gridDim = dim3(static_cast<unsigned>(units * CL));
    e = cudaLaunchKernelEx(&cfg, kernel, tx, tw, txh, a);
    if (e != cudaSuccess) { set_error("encoder_gemm: launch: %s", cudaGetErrorString(e)); return -2; }
    return 0;
}

}  // namespace enc

std::atomic<void*> g_enc_gemm_debug{nullptr};    // diagnostics: device buffer [grid][8] i64 of role timers
std::atomic<int> g_enc_gemm_form{0};     // 0 auto, 1 = <64, 1>, 2 = <256, 2> pairs, 3 = <256, 2> in clusters of four (X multicast)

int launch_encoder_gemm(const void* X, int64_t ldx, const void* W, const float* bias, int64_t m, int n, int k,
                        int epilogue, void* out0, int64_t ld0, void* out1, int64_t ld1, int n_split, int q_cols,
                        float q_scale, const float* residual, int64_t ldr, const float* res_stats,
                        const float* res_gamma, const float* res_beta, int sm_count, cudaStream_t stream) {
    using namespace enc;
    GemmArgs a = {};
    a.m = m;
    a.n = n;
    a.k = k;
    a.bias = bias;
    a.out0 = out0;
    a.ld0 = ld0;
    a.out1 = static_cast<__half*>(out1);
    a.ld1 = ld1;
    a.n_split = (epilogue == kEpiSplit) ? n_split : n;
    a.q_cols = q_cols;
    a.q_scale = q_scale;
    a.residual = residual;
    a.ldr = ldr;
    a.res_stats = reinterpret_cast<const float2*>(res_stats);
    a.res_gamma = res_gamma;
    a.res_beta = res_beta;
    a.dbg = static_cast<long long*>(g_enc_gemm_debug.load());
    if (m == 0) return 0;
    // pairs once there are enough 256 x 256 tiles to occupy most of the chip
    int form = g_enc_gemm_form.load();
    if (form == 0) form = (((m + 255) / 256) * (n / 256) >= sm_count / 4) ? 2 : 1;
    if (form == 4 && (n / 256) % 4 != 0) form = 3;            // clusters of eight take n-tiles four at a time
    if (form == 3 && (n / 256) % 2 != 0) form = 2;            // clusters of four take n-tiles two at a time
#define SQE_ENC_GEMM(EPI_)                                                                         \
    (form == 4 ? launch_gemm_t<256, 2, EPI_, 8>(X, W, a, ldx, sm_count, stream)                    \
     : form == 3 ? launch_gemm_t<256, 2, EPI_, 4>(X, W, a, ldx, sm_count, stream)                  \
     : form == 2 ? launch_gemm_t<256, 2, EPI_, 2>(X, W, a, ldx, sm_count, stream)                  \
                 : launch_gemm_t<64, 1, EPI_, 1>(X, W, a, ldx, sm_count, stream))
    switch (epilogue) {
        case kEpiSplit: return SQE_ENC_GEMM(kEpiSplit);
        case kEpiResF32: return SQE_ENC_GEMM(kEpiResF32);
        case kEpiGelu: return SQE_ENC_GEMM(kEpiGelu);
    }
#undef SQE_ENC_GEMM
    set_error("encoder_gemm: unknown epilogue %d", epilogue);
    return -1;
}

}  // namespace sqe
