// K1  fused L2-normalise + cast.
//
// Replaces, on the corpus side app/main.py:315-316 and app/embedding_gen.py:215-216
// and on the query side app/main.py:353-354 (paths relative to the reference):
//     norms = np.linalg.norm(E, axis=1, keepdims=True);  E = E / (norms + 1e-9)
//
// One warp owns one 1024-float row.  The row is read once with 8 coalesced 128-bit
// loads per lane (kept in registers for the divide) and mirrored to a padded shared
// tile so that each lane can walk the two strided accumulator chains numpy's
// pairwise summation assigns to it.  The summation tree below is numpy's fp32
// add.reduce order exactly (blocks of 128, eight stride-8 accumulators per block,
// ((r0+r1)+(r2+r3))+((r4+r5)+(r6+r7)), blocks combined by halving), the squares are
// rounded before they are added (no FMA contraction), sqrt and divide are the IEEE
// ones -- so the fp32 output is bit-identical to the reference expression, and the
// bf16/fp16 outputs are that value rounded to nearest even.
//
// HBM-bound: 4 KB read + 2/4 KB written per row; algorithmic bytes per row =
// 4096 + 1024*sizeof(out).
#include "sqe_common.cuh"
#include "sqe_internal.h"

namespace sqe {

constexpr int kNormWarps = 8;

__device__ __forceinline__ void store_row_chunk(float* out, float4 v) {
    asm volatile("st.global.L1::no_allocate.v4.f32 [%0], {%1,%2,%3,%4};" ::"l"(out), "f"(v.x),
                 "f"(v.y), "f"(v.z), "f"(v.w)
                 : "memory");
}
__device__ __forceinline__ void store_row_chunk(__nv_bfloat16* out, float4 v) {
    __nv_bfloat162 a = __floats2bfloat162_rn(v.x, v.y);
    __nv_bfloat162 b = __floats2bfloat162_rn(v.z, v.w);
    uint2 u;
    u.x = *reinterpret_cast<uint32_t*>(&a);
    u.y = *reinterpret_cast<uint32_t*>(&b);
    asm volatile("st.global.L1::no_allocate.v2.u32 [%0], {%1,%2};" ::"l"(out), "r"(u.x), "r"(u.y)
                 : "memory");
}
__device__ __forceinline__ void store_row_chunk(__half* out, float4 v) {
    __half2 a = __floats2half2_rn(v.x, v.y);
    __half2 b = __floats2half2_rn(v.z, v.w);
    uint2 u;
    u.x = *reinterpret_cast<uint32_t*>(&a);
    u.y = *reinterpret_cast<uint32_t*>(&b);
    asm volatile("st.global.L1::no_allocate.v2.u32 [%0], {%1,%2};" ::"l"(out), "r"(u.x), "r"(u.y)
                 : "memory");
}

// split bf16 (SQE_BF16X2): hi into the first 1024 bf16 of the row, lo into the second 1024
__device__ __forceinline__ void store_row_chunk(Bf16x2* out, float4 v) {
    __nv_bfloat16 h[4], l[4];
    split_bf16x2(v.x, h[0], l[0]);
    split_bf16x2(v.y, h[1], l[1]);
    split_bf16x2(v.z, h[2], l[2]);
    split_bf16x2(v.w, h[3], l[3]);
    uint2 uh, ul;
    uh.x = (static_cast<uint32_t>(__bfloat16_as_ushort(h[1])) << 16) | __bfloat16_as_ushort(h[0]);
    uh.y = (static_cast<uint32_t>(__bfloat16_as_ushort(h[3])) << 16) | __bfloat16_as_ushort(h[2]);
    ul.x = (static_cast<uint32_t>(__bfloat16_as_ushort(l[1])) << 16) | __bfloat16_as_ushort(l[0]);
    ul.y = (static_cast<uint32_t>(__bfloat16_as_ushort(l[3])) << 16) | __bfloat16_as_ushort(l[2]);
    asm volatile("st.global.L1::no_allocate.v2.u32 [%0], {%1,%2};" ::"l"(out), "r"(uh.x), "r"(uh.y) : "memory");
    asm volatile("st.global.L1::no_allocate.v2.u32 [%0], {%1,%2};" ::"l"(out + kDim), "r"(ul.x), "r"(ul.y) : "memory");
}
template <typename OutT> struct OutRow { static constexpr int kElems = kDim; };
template <> struct OutRow<Bf16x2> { static constexpr int kElems = 2 * kDim; };

template <typename OutT>
__global__ void __launch_bounds__(kNormWarps * 32, 4)
normalize_cast_kernel(const float* __restrict__ in, OutT* __restrict__ out, int64_t n) {
    __shared__ __align__(16) float tile[kNormWarps][8 * kNormBlockStride];
    const int lane = threadIdx.x & 31;
    const int warp = threadIdx.x >> 5;
    float* t = tile[warp];
    const int64_t warps_total = static_cast<int64_t>(gridDim.x) * kNormWarps;

    for (int64_t row = static_cast<int64_t>(blockIdx.x) * kNormWarps + warp; row < n;
         row += warps_total) {
        const float* src = in + row * kDim;
        float4 v[8];
#pragma unroll
        for (int m = 0; m < 8; ++m) {
            uint4 u = ldg_stream(src + 128 * m + 4 * lane);
            v[m] = make_float4(__uint_as_float(u.x), __uint_as_float(u.y), __uint_as_float(u.z),
                               __uint_as_float(u.w));
        }
        const float s = warp_row_sumsq_numpy(v, t, lane);
        const float den = __fadd_rn(__fsqrt_rn(s), 1e-9f);
        OutT* dst = out + row * OutRow<OutT>::kElems;
#pragma unroll
        for (int m = 0; m < 8; ++m) {
            float4 o;
            o.x = __fdiv_rn(v[m].x, den);
            o.y = __fdiv_rn(v[m].y, den);
            o.z = __fdiv_rn(v[m].z, den);
            o.w = __fdiv_rn(v[m].w, den);
            store_row_chunk(dst + 128 * m + 4 * lane, o);
        }
    }
}

int launch_normalize_cast(const float* in, void* out, int64_t n, int out_dtype, int sm_count,
                          cudaStream_t stream) {
    if (n == 0) return 0;
    int64_t blocks_needed = (n + kNormWarps - 1) / kNormWarps;
    int64_t grid = static_cast<int64_t>(sm_count) * 8;          // 2 waves of 4 resident CTAs per SM
    if (grid > blocks_needed) grid = blocks_needed;
    dim3 g(static_cast<unsigned>(grid)), b(kNormWarps * 32);
    switch (out_dtype) {
        case 0: normalize_cast_kernel<float><<<g, b, 0, stream>>>(in, static_cast<float*>(out), n); break;
        case 1: normalize_cast_kernel<__nv_bfloat16><<<g, b, 0, stream>>>(in, static_cast<__nv_bfloat16*>(out), n); break;
        case 2: normalize_cast_kernel<__half><<<g, b, 0, stream>>>(in, static_cast<__half*>(out), n); break;
        case 3: normalize_cast_kernel<Bf16x2><<<g, b, 0, stream>>>(in, static_cast<Bf16x2*>(out), n); break;
        default: return -1;
    }
    return 0;
}

}  // namespace sqe
