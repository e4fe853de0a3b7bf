// How one warp reads a stored row: per storage class, the 128-bit loads of a lane and the
// logical elements they hold.  Shared by the exact GEMV scan (topk_gemv.cu) and the prefiltered
// scan's exact rescoring pass (topk_prefilter.cu), which must score a row with the SAME fp32
// operation order so that both paths return bit-identical scores.
#pragma once
#include "sqe_common.cuh"

namespace sqe {

template <typename T> struct Elem;
template <> struct Elem<float> {
    static constexpr int kLoads = 8;     // 128-bit loads per lane per row
    static constexpr int kPer = 4;       // elements per load
    static constexpr int kQGroups = 8;   // groups of kPer logical elements per lane
    static constexpr int kRowElems = kDim;
    __device__ static __forceinline__ int load_off(int c) { return c * (32 * kPer); }
    __device__ static __forceinline__ void unpack(const uint4& u, float (&f)[4]) {
        f[0] = __uint_as_float(u.x); f[1] = __uint_as_float(u.y);
        f[2] = __uint_as_float(u.z); f[3] = __uint_as_float(u.w);
    }
};
template <> struct Elem<__nv_bfloat16> {
    static constexpr int kLoads = 4;
    static constexpr int kPer = 8;
    static constexpr int kQGroups = 4;
    static constexpr int kRowElems = kDim;
    __device__ static __forceinline__ int load_off(int c) { return c * (32 * kPer); }
    __device__ static __forceinline__ void unpack(const uint4& u, float (&f)[8]) {
        f[0] = __uint_as_float(u.x << 16); f[1] = __uint_as_float(u.x & 0xffff0000u);
        f[2] = __uint_as_float(u.y << 16); f[3] = __uint_as_float(u.y & 0xffff0000u);
        f[4] = __uint_as_float(u.z << 16); f[5] = __uint_as_float(u.z & 0xffff0000u);
        f[6] = __uint_as_float(u.w << 16); f[7] = __uint_as_float(u.w & 0xffff0000u);
    }
};
template <> struct Elem<__half> {
    static constexpr int kLoads = 4;
    static constexpr int kPer = 8;
    static constexpr int kQGroups = 4;
    static constexpr int kRowElems = kDim;
    __device__ static __forceinline__ int load_off(int c) { return c * (32 * kPer); }
    __device__ static __forceinline__ void unpack(const uint4& u, float (&f)[8]) {
        float2 a = __half22float2(*reinterpret_cast<const __half2*>(&u.x));
        float2 b = __half22float2(*reinterpret_cast<const __half2*>(&u.y));
        float2 c = __half22float2(*reinterpret_cast<const __half2*>(&u.z));
        float2 d = __half22float2(*reinterpret_cast<const __half2*>(&u.w));
        f[0] = a.x; f[1] = a.y; f[2] = b.x; f[3] = b.y;
        f[4] = c.x; f[5] = c.y; f[6] = d.x; f[7] = d.y;
    }
};

// split bf16: loads 0-3 fetch the hi plane, loads 4-7 the lo plane of the same 8-element
// groups; the value is hi + lo (exact in fp32)
template <> struct Elem<Bf16x2> {
    static constexpr int kLoads = 8;
    static constexpr int kPer = 8;
    static constexpr int kQGroups = 4;
    static constexpr int kRowElems = 2 * kDim;
    __device__ static __forceinline__ int load_off(int c) {
        return (c & 3) * (32 * kPer) + (c < 4 ? 0 : kDim);          // hi plane | lo plane
    }
    __device__ static __forceinline__ void unpack(const uint4& u, float (&f)[8]) {
        Elem<__nv_bfloat16>::unpack(u, f);
    }
};

// logical elements g*kPer .. +kPer-1 of a row held as raw 128-bit loads
template <typename T>
__device__ __forceinline__ void elem_group(const uint4* raw, int g, float (&f)[Elem<T>::kPer]) {
    Elem<T>::unpack(raw[g], f);
}
template <>
__device__ __forceinline__ void elem_group<Bf16x2>(const uint4* raw, int g, float (&f)[8]) {
    float lo[8];
    Elem<Bf16x2>::unpack(raw[g], f);
    Elem<Bf16x2>::unpack(raw[g + 4], lo);
#pragma unroll
    for (int e = 0; e < 8; ++e) f[e] = __fadd_rn(f[e], lo[e]);
}

// x rounded to the storage class and back to fp32: the value K1 would have stored
template <typename T> __device__ __forceinline__ float round_through(float x);
template <> __device__ __forceinline__ float round_through<Bf16x2>(float x) {
    __nv_bfloat16 hi, lo;
    split_bf16x2(x, hi, lo);
    return __fadd_rn(__bfloat162float(hi), __bfloat162float(lo));
}
template <> __device__ __forceinline__ float round_through<float>(float x) { return x; }
template <> __device__ __forceinline__ float round_through<__nv_bfloat16>(float x) {
    return __bfloat162float(__float2bfloat16_rn(x));
}
template <> __device__ __forceinline__ float round_through<__half>(float x) {
    return __half2float(__float2half_rn(x));
}

// The scans' fused query normalisation (called by ONE warp): s_q[0..1023] = the RAW fp32 query
// `src` normalised with the K1 arithmetic bit for bit (x / (|x| + 1e-9), app/main.py:353-354) and
// rounded to the shard's storage class -- what K1 followed by a load of the stored query gives.
// `s_tile` = 8 * kNormBlockStride floats of scratch.
template <typename T>
__device__ __forceinline__ void normalize_query_to_smem(const float* src, float* s_q, float* s_tile, int lane) {
    float4 v[8];
#pragma unroll
    for (int m = 0; m < 8; ++m) v[m] = *reinterpret_cast<const float4*>(src + 128 * m + 4 * lane);
    const float ss = warp_row_sumsq_numpy(v, s_tile, lane);
    const float den = __fadd_rn(__fsqrt_rn(ss), 1e-9f);
#pragma unroll
    for (int m = 0; m < 8; ++m) {
        float4 o;
        o.x = round_through<T>(__fdiv_rn(v[m].x, den));
        o.y = round_through<T>(__fdiv_rn(v[m].y, den));
        o.z = round_through<T>(__fdiv_rn(v[m].z, den));
        o.w = round_through<T>(__fdiv_rn(v[m].w, den));
        *reinterpret_cast<float4*>(s_q + 128 * m + 4 * lane) = o;
    }
}

}  // namespace sqe
