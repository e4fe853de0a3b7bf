// Shared device helpers: composite ordering keys, streaming loads, and the
// warp-resident sorted list used by every top-k stage (GEMV scan, CTA merge,
// grid merge, candidate compaction of the tensor-core path, shard merge).
//
// Ordering contract (include/sqe_b200.h): (score desc, row index asc).  A
// candidate is one 64-bit key
//        key = orderable_u32(score) << 32 | (0xFFFFFFFF - row)
// so "better" is simply "larger key" and a single max/min implements the
// reference's strict-'>' first-maximum rule (app/main.py:84) for any k.
// key == 0 is the empty slot (it is below every real key).
#pragma once
#include <cuda_runtime.h>
#include <cuda_bf16.h>
#include <cuda_fp16.h>
#include <stdint.h>

namespace sqe {

constexpr int kDim = 1024;
constexpr unsigned kFull = 0xffffffffu;

// Storage type SQE_BF16X2: every value x is stored as TWO bf16 numbers, hi = bf16(x) and
// lo = bf16(x - hi); a row is [hi[1024] | lo[1024]] = 4096 bytes (the size of an fp32 row).  The
// stored value is hi + lo, which is exact in fp32 (16 significant bits) -- the fp32 tolerance
// class on the bf16 tensor cores.  The tag type is 2 bytes so pointer arithmetic counts bf16s.
struct Bf16x2 {
    __nv_bfloat16 v;
};
__device__ __forceinline__ void split_bf16x2(float x, __nv_bfloat16& hi, __nv_bfloat16& lo) {
    hi = __float2bfloat16_rn(x);
    lo = __float2bfloat16_rn(__fsub_rn(x, __bfloat162float(hi)));
}

__device__ __forceinline__ uint32_t orderable_u32(float s) {
    s = s + 0.0f;                                   // -0.0 -> +0.0 (numpy treats them as equal)
    uint32_t u = __float_as_uint(s);
    return (u & 0x80000000u) ? ~u : (u | 0x80000000u);
}
__device__ __forceinline__ float from_orderable_u32(uint32_t o) {
    uint32_t u = (o & 0x80000000u) ? (o & 0x7fffffffu) : ~o;
    return __uint_as_float(u);
}
// A NaN score (a broken embedding) ranks below everything, like numpy's argsort puts NaN last
// and like the reference's `sim > best_sim` (app/main.py:84), which is never true for NaN.
__device__ __forceinline__ uint64_t make_key(float s, uint32_t row) {
    s = (s == s) ? s : __int_as_float(0xff800000);
    return (static_cast<uint64_t>(orderable_u32(s)) << 32) | static_cast<uint64_t>(0xffffffffu - row);
}
__device__ __forceinline__ float key_score(uint64_t key) {
    return from_orderable_u32(static_cast<uint32_t>(key >> 32));
}
__device__ __forceinline__ uint32_t key_row(uint64_t key) {
    return 0xffffffffu - static_cast<uint32_t>(key & 0xffffffffu);
}

__device__ __forceinline__ uint64_t shfl_u64(uint64_t v, int src) {
    uint32_t lo = __shfl_sync(kFull, static_cast<uint32_t>(v), src);
    uint32_t hi = __shfl_sync(kFull, static_cast<uint32_t>(v >> 32), src);
    return (static_cast<uint64_t>(hi) << 32) | lo;
}
__device__ __forceinline__ uint64_t shfl_xor_u64(uint64_t v, int mask) {
    uint32_t lo = __shfl_xor_sync(kFull, static_cast<uint32_t>(v), mask);
    uint32_t hi = __shfl_xor_sync(kFull, static_cast<uint32_t>(v >> 32), mask);
    return (static_cast<uint64_t>(hi) << 32) | lo;
}
__device__ __forceinline__ uint64_t shfl_up_u64(uint64_t v, int delta) {
    uint32_t lo = __shfl_up_sync(kFull, static_cast<uint32_t>(v), delta);
    uint32_t hi = __shfl_up_sync(kFull, static_cast<uint32_t>(v >> 32), delta);
    return (static_cast<uint64_t>(hi) << 32) | lo;
}
__device__ __forceinline__ uint64_t umax64(uint64_t a, uint64_t b) { return a > b ? a : b; }
__device__ __forceinline__ uint64_t umin64(uint64_t a, uint64_t b) { return a < b ? a : b; }

// 128-bit streaming load: read-only path, do not allocate in L1 (each shard
// byte is used once per query pass).
__device__ __forceinline__ uint4 ldg_stream(const void* p) {
    uint4 r;
    asm volatile("ld.global.nc.L1::no_allocate.v4.u32 {%0,%1,%2,%3}, [%4];"
                 : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w) : "l"(p));
    return r;
}

// ---------------------------------------------------------------------------
// WarpList<R>: L = 32*R keys held by one warp, element i = r*32 + lane, kept
// sorted descending (element 0 is the best).  All loops are unrolled so the
// register array is statically indexed.
// ---------------------------------------------------------------------------
template <int R>
struct WarpList {
    static constexpr int L = 32 * R;
    uint64_t key[R];

    __device__ __forceinline__ void clear() {
#pragma unroll
        for (int r = 0; r < R; ++r) key[r] = 0ull;
    }
    // worst key currently held (element L-1); warp-uniform
    __device__ __forceinline__ uint64_t worst() const { return shfl_u64(key[R - 1], 31); }

    // Insert one candidate known to every lane.  Caller has checked cand > worst().
    __device__ __forceinline__ void insert(uint64_t cand, int lane) {
#pragma unroll
        for (int r = R - 1; r >= 0; --r) {
            uint64_t up = shfl_up_u64(key[r], 1);             // element i-1 for lane > 0
            if (r > 0) {
                uint64_t wrap = shfl_u64(key[r - 1], 31);     // element i-1 for lane 0
                if (lane == 0) up = wrap;
            } else if (lane == 0) {
                up = ~0ull;
            }
            uint64_t mine = key[r];
            key[r] = (cand > mine) ? ((cand > up) ? up : cand) : mine;
        }
    }

    // In: a bitonic sequence.  Out: sorted descending.
    __device__ __forceinline__ void bitonic_merge(int lane) {
#pragma unroll
        for (int d = R / 2; d >= 1; d >>= 1) {                // element distance 32*d: in-lane
#pragma unroll
            for (int r = 0; r < R; ++r) {
                if ((r & d) == 0) {
                    uint64_t a = key[r], b = key[r + d];
                    key[r] = umax64(a, b);
                    key[r + d] = umin64(a, b);
                }
            }
        }
#pragma unroll
        for (int d = 16; d >= 1; d >>= 1) {                   // element distance d < 32: cross-lane
            const bool upper = (lane & d) != 0;
#pragma unroll
            for (int r = 0; r < R; ++r) {
                uint64_t o = shfl_xor_u64(key[r], d);
                key[r] = upper ? umin64(key[r], o) : umax64(key[r], o);
            }
        }
    }

    // Keep the best L of (this list, other list); both sorted descending.
    __device__ __forceinline__ void merge_sorted(const uint64_t (&other)[R], int lane) {
#pragma unroll
        for (int r = 0; r < R; ++r) {
            uint64_t rev = shfl_u64(other[R - 1 - r], 31 - lane);   // other[L-1-i]
            key[r] = umax64(key[r], rev);
        }
        bitonic_merge(lane);
    }

    // Sort arbitrary contents descending (full bitonic network).
    __device__ __forceinline__ void sort(int lane) {
#pragma unroll
        for (int s = 2; s <= L; s <<= 1) {
#pragma unroll
            for (int d = s >> 1; d >= 1; d >>= 1) {
                if (d >= 32) {
                    const int dr = d >> 5;
#pragma unroll
                    for (int r = 0; r < R; ++r) {
                        if ((r & dr) == 0) {
                            const bool desc = (((r * 32) & s) == 0) || (s == L);
                            uint64_t a = key[r], b = key[r + dr];
                            uint64_t hi = umax64(a, b), lo = umin64(a, b);
                            key[r] = desc ? hi : lo;
                            key[r + dr] = desc ? lo : hi;
                        }
                    }
                } else {
                    const bool lower = (lane & d) == 0;
#pragma unroll
                    for (int r = 0; r < R; ++r) {
                        const int i = r * 32 + lane;
                        const bool desc = ((i & s) == 0) || (s == L);
                        uint64_t o = shfl_xor_u64(key[r], d);
                        key[r] = (lower == desc) ? umax64(key[r], o) : umin64(key[r], o);
                    }
                }
            }
        }
    }

    // R == 1 only.  Sort descending when only elements 0..P-1 may be non-empty (P a power of
    // two): the bitonic stages above P would only shuffle empty slots.
    template <int P>
    __device__ __forceinline__ void sort_prefix(int lane) {
        static_assert(R == 1, "sort_prefix works on a 32-key list");
#pragma unroll
        for (int s = 2; s <= P; s <<= 1) {
#pragma unroll
            for (int d = s >> 1; d >= 1; d >>= 1) {
                const bool lower = (lane & d) == 0;
                const bool desc = ((lane & s) == 0) || (s == P);
                uint64_t o = shfl_xor_u64(key[0], d);
                key[0] = (lower == desc) ? umax64(key[0], o) : umin64(key[0], o);
            }
        }
    }

    __device__ __forceinline__ void load(const uint64_t* p, int lane) {
#pragma unroll
        for (int r = 0; r < R; ++r) key[r] = p[r * 32 + lane];
    }
    __device__ __forceinline__ void store(uint64_t* p, int lane) const {
#pragma unroll
        for (int r = 0; r < R; ++r) p[r * 32 + lane] = key[r];
    }
};

// ---------------------------------------------------------------------------
// Sum of squares of one 1024-float row in EXACTLY numpy's fp32 add.reduce order (pairwise
// summation: blocks of 128, eight stride-8 accumulators per block combined as
// ((r0+r1)+(r2+r3))+((r4+r5)+(r6+r7)), blocks combined by halving), squares rounded before
// they are added (no FMA contraction).  The warp holds the row as v[m] = elements
// 128 m + 4 lane .. +3; `t` is a per-warp smem tile of 8 * kNormBlockStride floats.  Used by
// K1 and by the fused query normalisation of K3, so both are bit-identical to
// np.linalg.norm(x)**2 of the reference expression (app/main.py:315-316, :353-354).
// ---------------------------------------------------------------------------
constexpr int kNormBlockStride = 136;        // 128 floats + 8 pad: conflict-free LDS.64 walk

__device__ __forceinline__ float warp_row_sumsq_numpy(const float4 (&v)[8], float* t, int lane) {
#pragma unroll
    for (int m = 0; m < 8; ++m)
        *reinterpret_cast<float4*>(t + kNormBlockStride * m + 4 * lane) = v[m];
    __syncwarp();
    const int blk = lane >> 2;               // which 128-block this lane sums
    const int jj = (lane & 3) * 2;           // accumulator pair (jj, jj+1) of that block
    const float* c = t + kNormBlockStride * blk + jj;
    float2 x0 = *reinterpret_cast<const float2*>(c);
    float r0 = __fmul_rn(x0.x, x0.x);
    float r1 = __fmul_rn(x0.y, x0.y);
#pragma unroll
    for (int i = 1; i < 16; ++i) {
        float2 x = *reinterpret_cast<const float2*>(c + 8 * i);
        r0 = __fadd_rn(r0, __fmul_rn(x.x, x.x));
        r1 = __fadd_rn(r1, __fmul_rn(x.y, x.y));
    }
    float s = __fadd_rn(r0, r1);                                   // (r0+r1) | (r2+r3) | ...
    s = __fadd_rn(s, __shfl_xor_sync(kFull, s, 1));                // pairs of pairs
    s = __fadd_rn(s, __shfl_xor_sync(kFull, s, 2));                // one 128-block
    s = __fadd_rn(s, __shfl_xor_sync(kFull, s, 4));                // 256
    s = __fadd_rn(s, __shfl_xor_sync(kFull, s, 8));                // 512
    s = __fadd_rn(s, __shfl_xor_sync(kFull, s, 16));               // 1024
    __syncwarp();                                                   // tile may be reused
    return s;
}

// Write the first k elements of a sorted list as (score, index) pairs.
template <int R>
__device__ __forceinline__ void emit_topk(const WarpList<R>& list, int k, int lane,
                                          float* out_score, int64_t* out_idx, int64_t idx_offset) {
#pragma unroll
    for (int r = 0; r < R; ++r) {
        const int i = r * 32 + lane;
        if (i < k) {
            const uint64_t key = list.key[r];
            if (key == 0ull) {
                out_score[i] = __int_as_float(0xff800000);      // -inf
                out_idx[i] = -1;
            } else {
                out_score[i] = key_score(key);
                out_idx[i] = idx_offset + static_cast<int64_t>(key_row(key));
            }
        }
    }
}

}  // namespace sqe
