// Row-wise encoder kernels (HBM-bound, one warp per 1024-element row, eight rows per CTA):
//   embed_ln   word + position + token-type embeddings -> LayerNorm -> fp32 row + fp16 copy
//   layernorm  fp32 pre-LayerNorm sum (written by the GEMM epilogue) -> fp32 row + fp16 copy
//   pool       CLS pooling: the first token of every sequence -> [n_seq, 1024] fp32, optionally
//              written straight into a query / ingest staging buffer of the retrieval path
// The fp32 row is the residual input of the next block, the fp16 copy its tensor-core operand.
// LayerNorm is the two-pass form (mean, then the variance of the deviations, biased, eps inside
// the square root) in fp32, as torch.nn.LayerNorm computes it.
#include "sqe_enc.cuh"

namespace sqe {
namespace enc {

constexpr int kRowWarps = 8;

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(kFull, v, o);
    return v;
}

// x: this lane's 32 values of the row, element index 4 * (lane + 32 * j) + e for x[4 * j + e]
// out32 may be null (the fp32 output is not needed: whoever needs it recomputes it from the input and
// `stats`); stats (may be null) receives {mean, rstd} of the row.
__device__ __forceinline__ void ln_row_store(float (&x)[32], const float* gamma, const float* beta, float eps,
                                             float* out32, __half* out16, float2* stats, int lane) {
    float s = 0.0f;
#pragma unroll
    for (int i = 0; i < 32; ++i) s += x[i];
    const float mean = warp_sum(s) * (1.0f / kHidden);
    float q = 0.0f;
#pragma unroll
    for (int i = 0; i < 32; ++i) {
        x[i] -= mean;
        q = fmaf(x[i], x[i], q);
    }
    const float rstd = rsqrtf(warp_sum(q) * (1.0f / kHidden) + eps);
    if (stats != nullptr && lane == 0) *stats = make_float2(mean, rstd);
#pragma unroll
    for (int j = 0; j < 8; ++j) {
        const int c = 4 * (lane + 32 * j);
        const float4 g = __ldg(reinterpret_cast<const float4*>(gamma + c));
        const float4 b = __ldg(reinterpret_cast<const float4*>(beta + c));
        float4 y;
        y.x = fmaf(x[4 * j + 0] * rstd, g.x, b.x);
        y.y = fmaf(x[4 * j + 1] * rstd, g.y, b.y);
        y.z = fmaf(x[4 * j + 2] * rstd, g.z, b.z);
        y.w = fmaf(x[4 * j + 3] * rstd, g.w, b.w);
        if (out32 != nullptr) *reinterpret_cast<float4*>(out32 + c) = y;
        const __half2 h0 = __floats2half2_rn(y.x, y.y), h1 = __floats2half2_rn(y.z, y.w);
        uint2 u;
        u.x = *reinterpret_cast<const uint32_t*>(&h0);
        u.y = *reinterpret_cast<const uint32_t*>(&h1);
        *reinterpret_cast<uint2*>(out16 + c) = u;
    }
}

__global__ void __launch_bounds__(kRowWarps * 32)
encoder_layernorm_kernel(const float* __restrict__ in, const float* __restrict__ gamma, const float* __restrict__ beta,
                         float eps, int64_t rows, float* __restrict__ out32, __half* __restrict__ out16,
                         float2* __restrict__ stats) {
    const int lane = threadIdx.x & 31;
    const int64_t row = static_cast<int64_t>(blockIdx.x) * kRowWarps + (threadIdx.x >> 5);
    if (row >= rows) return;
    float x[32];
    const float4* ip = reinterpret_cast<const float4*>(in + row * kHidden);
#pragma unroll
    for (int j = 0; j < 8; ++j) {
        // with statistics the input is read again (as a residual, by the next GEMM's epilogue): keep it cached
        const float4 v = stats != nullptr ? __ldg(ip + lane + 32 * j) : __ldcs(ip + lane + 32 * j);
        x[4 * j + 0] = v.x; x[4 * j + 1] = v.y; x[4 * j + 2] = v.z; x[4 * j + 3] = v.w;
    }
    ln_row_store(x, gamma, beta, eps, out32 != nullptr ? out32 + row * kHidden : nullptr, out16 + row * kHidden,
                 stats != nullptr ? stats + row : nullptr, lane);
}

// ids[t] < 0 marks a padding row: it is written as zeros (finite keys for the attention kernel).
__global__ void __launch_bounds__(kRowWarps * 32)
encoder_embed_ln_kernel(const int32_t* __restrict__ ids, const int32_t* __restrict__ pos,
                        const float* __restrict__ word, const float* __restrict__ position,
                        const float* __restrict__ type0, const float* __restrict__ gamma,
                        const float* __restrict__ beta, float eps, int64_t rows, int vocab, int max_pos,
                        float* __restrict__ out32, __half* __restrict__ out16, float2* __restrict__ stats) {
    const int lane = threadIdx.x & 31;
    const int64_t row = static_cast<int64_t>(blockIdx.x) * kRowWarps + (threadIdx.x >> 5);
    if (row >= rows) return;
    const int id = __ldg(ids + row);
    float* o32 = out32 + row * kHidden;
    __half* o16 = out16 + row * kHidden;
    if (id < 0 || id >= vocab) {
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            const int c = 4 * (lane + 32 * j);
            *reinterpret_cast<float4*>(o32 + c) = make_float4(0.f, 0.f, 0.f, 0.f);
            *reinterpret_cast<uint2*>(o16 + c) = make_uint2(0u, 0u);
        }
        if (stats != nullptr && lane == 0) stats[row] = make_float2(0.f, 0.f);
        return;
    }
    int p = __ldg(pos + row);
    p = p < 0 ? 0 : (p >= max_pos ? max_pos - 1 : p);
    const float4* wp = reinterpret_cast<const float4*>(word + static_cast<int64_t>(id) * kHidden);
    const float4* pp = reinterpret_cast<const float4*>(position + static_cast<int64_t>(p) * kHidden);
    const float4* tp = reinterpret_cast<const float4*>(type0);
    float x[32];
#pragma unroll
    for (int j = 0; j < 8; ++j) {
        const float4 w = __ldg(wp + lane + 32 * j), q = __ldg(pp + lane + 32 * j), t = __ldg(tp + lane + 32 * j);
        // (word + token_type) + position: the order of BertEmbeddings.forward
        x[4 * j + 0] = (w.x + t.x) + q.x;
        x[4 * j + 1] = (w.y + t.y) + q.y;
        x[4 * j + 2] = (w.z + t.z) + q.z;
        x[4 * j + 3] = (w.w + t.w) + q.w;
    }
    if (stats != nullptr) {
        // statistics form: out32 receives the PRE-LayerNorm sum (the residual source of the first block)
#pragma unroll
        for (int j = 0; j < 8; ++j)
            *reinterpret_cast<float4*>(o32 + 4 * (lane + 32 * j)) = make_float4(x[4 * j], x[4 * j + 1], x[4 * j + 2], x[4 * j + 3]);
        ln_row_store(x, gamma, beta, eps, nullptr, o16, stats + row, lane);
    } else {
        ln_row_store(x, gamma, beta, eps, o32, o16, nullptr, lane);
    }
}

// out[s, :] = h[first_token[s], :] (fp32); one warp per sequence
// stats != null: h holds pre-LayerNorm sums; the pooled row is LayerNorm(h row) with these statistics
__global__ void __launch_bounds__(kRowWarps * 32)
encoder_pool_kernel(const float* __restrict__ h, const int32_t* __restrict__ first_token, int n_seq,
                    float* __restrict__ out, int64_t ldo, const float2* __restrict__ stats,
                    const float* __restrict__ gamma, const float* __restrict__ beta) {
    const int lane = threadIdx.x & 31;
    const int s = blockIdx.x * kRowWarps + (threadIdx.x >> 5);
    if (s >= n_seq) return;
    const int64_t row = __ldg(first_token + s);
    const float4* ip = reinterpret_cast<const float4*>(h + row * kHidden);
    float4* op = reinterpret_cast<float4*>(out + static_cast<int64_t>(s) * ldo);
    const float2 st = stats != nullptr ? __ldg(stats + row) : make_float2(0.f, 1.f);
#pragma unroll
    for (int j = 0; j < 8; ++j) {
        float4 v = __ldg(ip + lane + 32 * j);
        if (stats != nullptr) {
            const float4 g = __ldg(reinterpret_cast<const float4*>(gamma) + lane + 32 * j);
            const float4 b = __ldg(reinterpret_cast<const float4*>(beta) + lane + 32 * j);
            v.x = fmaf((v.x - st.x) * st.y, g.x, b.x);
            v.y = fmaf((v.y - st.x) * st.y, g.y, b.y);
            v.z = fmaf((v.z - st.x) * st.y, g.z, b.z);
            v.w = fmaf((v.w - st.x) * st.y, g.w, b.w);
        }
        op[lane + 32 * j] = v;
    }
}

}  // namespace enc

int launch_encoder_layernorm(const float* in, const float* gamma, const float* beta, float eps, int64_t rows,
                             float* out32, void* out16, float* stats, cudaStream_t stream) {
    using namespace enc;
    if (rows == 0) return 0;
    const unsigned grid = static_cast<unsigned>((rows + kRowWarps - 1) / kRowWarps);
    encoder_layernorm_kernel<<<grid, kRowWarps * 32, 0, stream>>>(in, gamma, beta, eps, rows, out32,
                                                                  static_cast<__half*>(out16),
                                                                  reinterpret_cast<float2*>(stats));
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) { set_error("encoder_layernorm: launch: %s", cudaGetErrorString(e)); return -2; }
    return 0;
}

int launch_encoder_embed_ln(const int32_t* ids, const int32_t* pos, const float* word, const float* position,
                            const float* type0, const float* gamma, const float* beta, float eps, int64_t rows,
                            int vocab, int max_pos, float* out32, void* out16, float* stats, cudaStream_t stream) {
    using namespace enc;
    if (rows == 0) return 0;
    const unsigned grid = static_cast<unsigned>((rows + kRowWarps - 1) / kRowWarps);
    encoder_embed_ln_kernel<<<grid, kRowWarps * 32, 0, stream>>>(ids, pos, word, position, type0, gamma, beta, eps,
                                                                 rows, vocab, max_pos, out32,
                                                                 static_cast<__half*>(out16),
                                                                 reinterpret_cast<float2*>(stats));
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) { set_error("encoder_embed_ln: launch: %s", cudaGetErrorString(e)); return -2; }
    return 0;
}

int launch_encoder_pool(const float* h, const int32_t* first_token, int n_seq, float* out, int64_t ldo,
                        const float* stats, const float* gamma, const float* beta, cudaStream_t stream) {
    using namespace enc;
    if (n_seq == 0) return 0;
    encoder_pool_kernel<<<(n_seq + kRowWarps - 1) / kRowWarps, kRowWarps * 32, 0, stream>>>(
        h, first_token, n_seq, out, ldo, reinterpret_cast<const float2*>(stats), gamma, beta);
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) { set_error("encoder_pool: launch: %s", cudaGetErrorString(e)); return -2; }
    return 0;
}

}  // namespace sqe
