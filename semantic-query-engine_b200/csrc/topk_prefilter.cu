// K3p  batch-1 exact cosine top-k at HALF the HBM traffic: an int8 prefilter with a rigorous
// error bound, followed by exact rescoring of the few rows the bound cannot rule out.
//
// Same contract and same results as K3 (topk_gemv.cu) -- which replaces the k-NN request of
// OpenSearchIndexer.search, app/main.py:356-367 -- bit for bit: scores, rows, tie order.  K3 is
// HBM-bound (it streams every stored row, 2 KB for bf16); the only way past that roofline is to
// read fewer bytes.  Next to the shard the index keeps a COARSE copy: every row quantised to
// int8 with its own scale (1 KB per row) plus 16 bytes of row constants.
//
//   ingest   quantize_rows_kernel (K1q):  d8 = rint(d / sd), sd = max|d| / 127,
//            eps = an upper bound of |d - sd d8|_2, nd = an upper bound of |sd d8|_2
//   pass A   coarse_scan_kernel: the query is quantised the same way (q = sq q8 + eq); every row
//            costs 8 dp4a per lane (exact integer arithmetic): s8 = sd sq (q8 . d8).  By
//            Cauchy-Schwarz   | q.d - s8 | <= |eq| nd + |q| eps   =: m   (q.d = the real-number
//            dot product of the stored values), so  L = s8 - m - f  <=  score  <=  s8 + m + f = U,
//            where `score` is the fp32 number K3 computes and f covers its rounding.  The kernel
//            writes U for every row (4 B per row) and selects the top-k of L exactly like K3
//            selects scores; tau = the k-th best L is a lower bound of the k-th best score.
//   pass B   rescore_kernel: scans U; a row with U < tau cannot be in the result (score <= U <
//            tau <= k-th best score); every other row is scored with K3's own arithmetic (same
//            loads, same FMA chains, same butterfly -> the same bits) and goes through the same
//            warp / CTA / last-CTA selection.  For unit-norm embedding rows m ~ 0.017 while the
//            score spread is sigma = 1/32, so ~1e-4 of the rows are rescored.
//
// Nothing is approximate: if the bound is loose (clustered data, non-finite values, fewer than k
// rows: tau = -inf) more rows are rescored, in the limit all of them, and the result is still
// K3's.  Non-finite stored rows get eps = +inf (always rescored); a non-finite query makes every
// margin NaN and `!(U < tau)` sends every row to the exact pass.
//
// Roofline: HBM.  Algorithmic bytes per query = n * (1024 + 16 + 4 + 4): the int8 row, its
// constants, U written once and read once (+ 2 KB per rescored row).  10M rows: 10.5 GB instead
// of K3's 20.5 GB.
#include "sqe_common.cuh"
#include "sqe_internal.h"
#include "sqe_rowload.cuh"
#include "sqe_select.cuh"

#include <cstring>

namespace sqe {

namespace pf {
constexpr int kWarps = 8;
constexpr int kCtasPerSm = 2;
constexpr int kRowsPerIter = 8;            // int8 rows in flight per warp (2 x 128-bit loads each)
constexpr int kMaxQueries = 64;
// absolute slack per unit of |q| (|d| + eps): rounding of K3's fp32 score (<= 22 roundings on any
// path of its summation tree, 22 * 2^-24 = 1.3e-6) + rounding of s8 and of the margin itself
constexpr float kSlack = 4e-6f;
constexpr float kInflate = 1.001f;         // covers the fp32 rounding of the norms' summations

// workspace layout (bytes)
constexpr int64_t kOffCounters = 0;        // [0, 4096): ticket counters, zero between calls
constexpr int64_t kOffTau = 4096;          // float [64]
constexpr int64_t kOffStats = 4096 + 256;  // u32 [64]: rows rescored
constexpr int64_t kOffLists = 8192;
}  // namespace pf

// ------------------------------------------------------------------------------------------
// K1q: stored rows (any storage class) -> int8 rows + row constants {sd, eps, nd, 0}
// ------------------------------------------------------------------------------------------
template <typename T>
__global__ void __launch_bounds__(pf::kWarps * 32, 4)
quantize_rows_kernel(const T* __restrict__ D, int64_t n, int8_t* __restrict__ D8,
                     float4* __restrict__ meta) {
    using E = Elem<T>;
    constexpr int LOADS = E::kLoads;
    constexpr int PER = E::kPer;
    constexpr int QG = E::kQGroups;
    const int lane = threadIdx.x & 31;
    const int warp = threadIdx.x >> 5;
    const int64_t warps_total = static_cast<int64_t>(gridDim.x) * pf::kWarps;
    for (int64_t row = static_cast<int64_t>(blockIdx.x) * pf::kWarps + warp; row < n; row += warps_total) {
        const T* rp = D + row * E::kRowElems + lane * PER;
        uint4 raw[LOADS];
#pragma unroll
        for (int c = 0; c < LOADS; ++c) raw[c] = ldg_stream(rp + E::load_off(c));
        float f[QG][PER];
        float mx = 0.f, ss = 0.f;
#pragma unroll
        for (int g = 0; g < QG; ++g) {
            elem_group<T>(raw, g, f[g]);
#pragma unroll
            for (int e = 0; e < PER; ++e) {
                mx = fmaxf(mx, fabsf(f[g][e]));
                ss = fmaf(f[g][e], f[g][e], ss);
            }
        }
#pragma unroll
        for (int d = 16; d >= 1; d >>= 1) {
            mx = fmaxf(mx, __shfl_xor_sync(kFull, mx, d));
            ss += __shfl_xor_sync(kFull, ss, d);
        }
        const bool finite = (ss - ss) == 0.f;                  // false for inf and NaN
        const bool live = finite && mx > 0.f;
        const float sd = live ? mx / 127.f : 0.f;
        const float inv = live ? 127.f / mx : 0.f;
        float e2 = 0.f, n2 = 0.f;
        int8_t* dst = D8 + row * kDim;
#pragma unroll
        for (int g = 0; g < QG; ++g) {
            uint32_t packed[PER / 4];
#pragma unroll
            for (int w = 0; w < PER / 4; ++w) packed[w] = 0u;
#pragma unroll
            for (int e = 0; e < PER; ++e) {
                const float x = live ? f[g][e] : 0.f;
                float r = rintf(x * inv);
                r = fminf(fmaxf(r, -127.f), 127.f);
                const float back = sd * r;
                const float err = fmaf(-sd, r, x);
                e2 = fmaf(err, err, e2);
                n2 = fmaf(back, back, n2);
                const int qi = static_cast<int>(r);
                packed[e >> 2] |= (static_cast<uint32_t>(qi) & 0xffu) << (8 * (e & 3));
            }
            int8_t* p = dst + g * (32 * PER) + lane * PER;
            if constexpr (PER == 8) {
                *reinterpret_cast<uint2*>(p) = make_uint2(packed[0], packed[1]);
            } else {
                *reinterpret_cast<uint32_t*>(p) = packed[0];
            }
        }
#pragma unroll
        for (int d = 16; d >= 1; d >>= 1) {
            e2 += __shfl_xor_sync(kFull, e2, d);
            n2 += __shfl_xor_sync(kFull, n2, d);
        }
        if (lane == 0) {
            float4 m;
            m.x = sd;
            m.y = finite ? sqrtf(e2) * pf::kInflate + 1e-12f : __int_as_float(0x7f800000);
            m.z = finite ? sqrtf(n2) * pf::kInflate : 0.f;
            m.w = 0.f;
            meta[row] = m;
        }
    }
}

// value of element i of a stored query row as fp32 (hi + lo for split bf16)
template <typename T> __device__ __forceinline__ float stored_elem(const T* q, int i);
template <> __device__ __forceinline__ float stored_elem<float>(const float* q, int i) { return q[i]; }
template <> __device__ __forceinline__ float stored_elem<__nv_bfloat16>(const __nv_bfloat16* q, int i) {
    return __bfloat162float(q[i]);
}
template <> __device__ __forceinline__ float stored_elem<__half>(const __half* q, int i) { return __half2float(q[i]); }
template <> __device__ __forceinline__ float stored_elem<Bf16x2>(const Bf16x2* q, int i) {
    return __fadd_rn(__bfloat162float(q[i].v), __bfloat162float(q[i + kDim].v));
}

__device__ __forceinline__ float block_reduce(float v, float* s_red, bool is_max, int warp, int lane) {
#pragma unroll
    for (int d = 16; d >= 1; d >>= 1) {
        const float o = __shfl_xor_sync(kFull, v, d);
        v = is_max ? fmaxf(v, o) : v + o;
    }
    __syncthreads();                                             // s_red may still be read
    if (lane == 0) s_red[warp] = v;
    __syncthreads();
    float r = s_red[0];
#pragma unroll
    for (int w = 1; w < pf::kWarps; ++w) r = is_max ? fmaxf(r, s_red[w]) : r + s_red[w];
    return r;
}

// ------------------------------------------------------------------------------------------
// pass A: int8 scan.  grid = (queries, CTAs per query), query fast (rows shared through L2).
// ------------------------------------------------------------------------------------------
// RAWQ: `Qv` holds RAW fp32 queries; every CTA normalises its query itself (K1's arithmetic bit for
// bit, sqe_rowload.cuh) -- the K1 launch in front of a one-query search disappears.
template <typename T, int R, bool RAWQ>
__global__ void __launch_bounds__(pf::kWarps * 32, pf::kCtasPerSm)
coarse_scan_kernel(const int8_t* __restrict__ D8, const float4* __restrict__ meta, int64_t n,
                   const void* __restrict__ Qv, int k, uint64_t* __restrict__ ws_lists,
                   unsigned* __restrict__ ws_counter, float* __restrict__ ws_tau,
                   unsigned* __restrict__ ws_stats, float* __restrict__ U) {
    constexpr int L = 32 * R;
    constexpr int RPI = pf::kRowsPerIter;
    __shared__ uint64_t s_lists[pf::kWarps][L];
    __shared__ int s_is_last;
    __shared__ __align__(16) int8_t s_q8[kDim];
    __shared__ float s_red[pf::kWarps];

    const int lane = threadIdx.x & 31;
    const int warp = threadIdx.x >> 5;
    const int query = blockIdx.x;
    const int cta = blockIdx.y;
    const int nctas = gridDim.y;

    // ---- quantise the query: q = sq q8 + eq ----
    [[maybe_unused]] __shared__ __align__(16) float s_qf[RAWQ ? kDim : 4];
    [[maybe_unused]] __shared__ __align__(16) float s_tile[RAWQ ? 8 * kNormBlockStride : 4];
    if constexpr (RAWQ) {
        if (warp == 0)
            normalize_query_to_smem<T>(static_cast<const float*>(Qv) + static_cast<int64_t>(query) * kDim, s_qf,
                                       s_tile, lane);
        __syncthreads();
    }
    const T* qrow = static_cast<const T*>(Qv) + static_cast<int64_t>(query) * Elem<T>::kRowElems;
    float qv[kDim / (pf::kWarps * 32)];
    float mx = 0.f, ss = 0.f;
#pragma unroll
    for (int j = 0; j < kDim / (pf::kWarps * 32); ++j) {
        if constexpr (RAWQ) qv[j] = s_qf[j * (pf::kWarps * 32) + threadIdx.x];
        else qv[j] = stored_elem<T>(qrow, j * (pf::kWarps * 32) + threadIdx.x);
        mx = fmaxf(mx, fabsf(qv[j]));
        ss = fmaf(qv[j], qv[j], ss);
    }
    mx = block_reduce(mx, s_red, true, warp, lane);
    ss = block_reduce(ss, s_red, false, warp, lane);
    const bool live = mx > 0.f && (mx - mx) == 0.f;
    const float sq = live ? mx / 127.f : 0.f;
    const float inv = live ? 127.f / mx : 0.f;
    float e2 = 0.f;
#pragma unroll
    for (int j = 0; j < kDim / (pf::kWarps * 32); ++j) {
        float r = rintf(qv[j] * inv);
        r = fminf(fmaxf(r, -127.f), 127.f);
        r = (r == r) ? r : 0.f;                                  // a NaN element: ss is NaN as well
        const float err = fmaf(-sq, r, qv[j]);
        e2 = fmaf(err, err, e2);
        s_q8[j * (pf::kWarps * 32) + threadIdx.x] = static_cast<int8_t>(static_cast<int>(r));
    }
    e2 = block_reduce(e2, s_red, false, warp, lane);
    // |eq| and |q| (upper bounds); NaN / inf in the query make both NaN / inf -> every row is rescored
    const float qe = sqrtf(e2) * pf::kInflate + 1e-12f;
    const float qn = sqrtf(ss) * pf::kInflate;
    __syncthreads();
    int q8[8];
#pragma unroll
    for (int c = 0; c < 2; ++c) {
        const int4 v = *reinterpret_cast<const int4*>(s_q8 + c * 512 + lane * 16);
        q8[4 * c + 0] = v.x; q8[4 * c + 1] = v.y; q8[4 * c + 2] = v.z; q8[4 * c + 3] = v.w;
    }

    WarpList<R> list;
    list.clear();
    uint64_t worst = 0ull;
    float* Uq = U + static_cast<int64_t>(query) * n;

    const int64_t warps_total = static_cast<int64_t>(nctas) * pf::kWarps;
    const int64_t gw = static_cast<int64_t>(cta) * pf::kWarps + warp;
    for (int64_t base = gw * RPI; base < n; base += warps_total * RPI) {
        uint4 raw[RPI][2];
#pragma unroll
        for (int j = 0; j < RPI; ++j) {
            const int64_t row = base + j;
            if (row < n) {
                const int8_t* rp = D8 + row * kDim + lane * 16;
                raw[j][0] = ldg_stream(rp);
                raw[j][1] = ldg_stream(rp + 512);
            } else {
                raw[j][0] = make_uint4(0, 0, 0, 0);
                raw[j][1] = make_uint4(0, 0, 0, 0);
            }
        }
        // lane j (< 8) owns row base + j from here on
        const int64_t myrow = base + lane;
        const bool mine = lane < RPI && myrow < n;
        float4 mt = make_float4(0.f, 0.f, 0.f, 0.f);
        if (mine) {
            const uint4 u = ldg_stream(meta + myrow);
            mt = make_float4(__uint_as_float(u.x), __uint_as_float(u.y), __uint_as_float(u.z), __uint_as_float(u.w));
        }
        int acc[RPI];
#pragma unroll
        for (int j = 0; j < RPI; ++j) {
            int a = 0;
            a = __dp4a(static_cast<int>(raw[j][0].x), q8[0], a);
            a = __dp4a(static_cast<int>(raw[j][0].y), q8[1], a);
            a = __dp4a(static_cast<int>(raw[j][0].z), q8[2], a);
            a = __dp4a(static_cast<int>(raw[j][0].w), q8[3], a);
            a = __dp4a(static_cast<int>(raw[j][1].x), q8[4], a);
            a = __dp4a(static_cast<int>(raw[j][1].y), q8[5], a);
            a = __dp4a(static_cast<int>(raw[j][1].z), q8[6], a);
            a = __dp4a(static_cast<int>(raw[j][1].w), q8[7], a);
            acc[j] = a;
        }
#pragma unroll
        for (int d = 16; d >= 1; d >>= 1) {
#pragma unroll
            for (int j = 0; j < RPI; ++j) acc[j] += __shfl_xor_sync(kFull, acc[j], d);
        }
        int a = acc[0];
#pragma unroll
        for (int j = 1; j < RPI; ++j) a = (lane == j) ? acc[j] : a;
        uint64_t key = 0ull;
        if (mine) {
            const float s8 = (mt.x * sq) * static_cast<float>(a);
            const float m = (fmaf(qe, mt.z, qn * mt.y) + pf::kSlack * qn * (mt.z + mt.y)) * pf::kInflate + 1e-30f;
            Uq[myrow] = s8 + m;
            key = make_key(s8 - m, static_cast<uint32_t>(myrow));     // NaN -> -inf
        }
        unsigned pend = __ballot_sync(kFull, key > worst);
        while (pend) {
            const int j = __ffs(pend) - 1;
            pend &= pend - 1;
            const uint64_t cand = shfl_u64(key, j);
            if (cand > worst) {
                list.insert(cand, lane);
                worst = list.worst();
            }
        }
    }

    pdl_launch_dependents();                       // the rescoring pass may be scheduled (it waits at its top)
    pdl_wait();                                    // lists / tickets / tau are shared with earlier kernels
    if (!merge_cta_and_grid<R, pf::kWarps>(list, s_lists, &s_is_last, ws_lists, ws_counter + query, query, cta, nctas, warp, lane))
        return;
    // last CTA, warp 0: tau = the k-th best lower bound (-inf while fewer than k rows exist)
    uint64_t kth_src = 0ull;
#pragma unroll
    for (int r = 0; r < R; ++r)
        if (r == ((k - 1) >> 5)) kth_src = list.key[r];
    const uint64_t kth = shfl_u64(kth_src, (k - 1) & 31);
    if (lane == 0) {
        ws_tau[query] = kth ? key_score(kth) : __int_as_float(0xff800000);
        ws_stats[query] = 0u;
        ws_counter[query] = 0u;
    }
}

// ------------------------------------------------------------------------------------------
// pass B: rescore every row the bound cannot rule out with K3's arithmetic, select the top-k
// ------------------------------------------------------------------------------------------
template <typename T, int R, bool RAWQ>
__global__ void __launch_bounds__(pf::kWarps * 32, pf::kCtasPerSm)
rescore_kernel(const T* __restrict__ D, int64_t n, const void* __restrict__ Qv, int k,
               const float* __restrict__ U, const float* __restrict__ ws_tau,
               unsigned* __restrict__ ws_stats, uint64_t* __restrict__ ws_lists,
               unsigned* __restrict__ ws_counter, float* __restrict__ out_score,
               int64_t* __restrict__ out_idx, int64_t idx_offset, unsigned* __restrict__ out_stats,
               const XchgArgs xchg) {
    using E = Elem<T>;
    constexpr int LOADS = E::kLoads;
    constexpr int PER = E::kPer;
    constexpr int QG = E::kQGroups;
    constexpr int L = 32 * R;
    constexpr int UNR = 8;                                       // 32-row strips of U in flight per warp
    __shared__ uint64_t s_lists[pf::kWarps][L];
    __shared__ int s_is_last;

    const int lane = threadIdx.x & 31;
    const int warp = threadIdx.x >> 5;
    const int query = blockIdx.x;
    const int cta = blockIdx.y;
    const int nctas = gridDim.y;

    // the query in registers, the element layout of a row's loads (= K3, stored-query form)
    float q[32];
    if constexpr (RAWQ) {
        __shared__ __align__(16) float s_tile[8 * kNormBlockStride];
        __shared__ __align__(16) float s_q[kDim];
        if (warp == 0)
            normalize_query_to_smem<T>(static_cast<const float*>(Qv) + static_cast<int64_t>(query) * kDim, s_q,
                                       s_tile, lane);
        __syncthreads();
#pragma unroll
        for (int g = 0; g < QG; ++g)
#pragma unroll
            for (int e = 0; e < PER; ++e) q[g * PER + e] = s_q[g * (32 * PER) + lane * PER + e];
    } else {
        const T* qp = static_cast<const T*>(Qv) + static_cast<int64_t>(query) * E::kRowElems + lane * PER;
        uint4 qraw[LOADS];
#pragma unroll
        for (int c = 0; c < LOADS; ++c) qraw[c] = *reinterpret_cast<const uint4*>(qp + E::load_off(c));
#pragma unroll
        for (int g = 0; g < QG; ++g) {
            float f[PER];
            elem_group<T>(qraw, g, f);
#pragma unroll
            for (int e = 0; e < PER; ++e) q[g * PER + e] = f[e];
        }
    }
    pdl_wait();                                    // tau and U come from the scan pass in front of this one
    const float tau = ws_tau[query];
    const float* Uq = U + static_cast<int64_t>(query) * n;

    WarpList<R> list;
    list.clear();
    uint64_t worst = 0ull;
    unsigned rescored = 0u;

    const int64_t warps_total = static_cast<int64_t>(nctas) * pf::kWarps;
    const int64_t gw = static_cast<int64_t>(cta) * pf::kWarps + warp;
    for (int64_t cb = gw * (32 * UNR); cb < n; cb += warps_total * (32 * UNR)) {
        float u[UNR];
#pragma unroll
        for (int t = 0; t < UNR; ++t) {
            const int64_t r = cb + t * 32 + lane;
            u[t] = (r < n) ? __ldcs(Uq + r) : 0.f;
        }
        // which (strip, lane) positions survive: one 32-bit mask per strip, then ONE loop over all
        // survivors (the strips are statically indexed; survivors are ~1e-4 of the rows)
        unsigned mask[UNR];
        unsigned any = 0u;
#pragma unroll
        for (int t = 0; t < UNR; ++t) {
            mask[t] = __ballot_sync(kFull, (cb + t * 32 + lane < n) && !(u[t] < tau));
            any |= mask[t];
            rescored += __popc(mask[t]);
        }
        while (any) {
            // lowest strip with a survivor, lowest lane in it (row order: ascending)
            int t = 0;
            unsigned hits = 0u;
#pragma unroll
            for (int tt = UNR - 1; tt >= 0; --tt)
                if (mask[tt]) { t = tt; hits = mask[tt]; }
            const int j = __ffs(hits) - 1;
#pragma unroll
            for (int tt = 0; tt < UNR; ++tt)
                if (tt == t) mask[tt] &= mask[tt] - 1;
            any = 0u;
#pragma unroll
            for (int tt = 0; tt < UNR; ++tt) any |= mask[tt];
            {
                const int64_t row = cb + t * 32 + j;
                const T* rp = D + row * E::kRowElems + lane * PER;
                uint4 raw[LOADS];
#pragma unroll
                for (int c = 0; c < LOADS; ++c) raw[c] = ldg_stream(rp + E::load_off(c));
                // K3's arithmetic, operation for operation (topk_gemv.cu)
                float a0 = 0.f, a1 = 0.f;
#pragma unroll
                for (int g = 0; g < QG; ++g) {
                    float f[PER];
                    elem_group<T>(raw, g, f);
#pragma unroll
                    for (int e = 0; e < PER; e += 2) {
                        a0 = fmaf(f[e], q[g * PER + e], a0);
                        a1 = fmaf(f[e + 1], q[g * PER + e + 1], a1);
                    }
                }
                float s = a0 + a1;
#pragma unroll
                for (int d = 16; d >= 1; d >>= 1) s += __shfl_xor_sync(kFull, s, d);
                const uint64_t key = make_key(s, static_cast<uint32_t>(row));
                if (key > worst) {
                    list.insert(key, lane);
                    worst = list.worst();
                }
            }
        }
    }
    if (lane == 0 && rescored) atomicAdd(ws_stats + query, rescored);
    pdl_launch_dependents();                       // U has been read: the next query's scan pass may overwrite it

    if (!merge_cta_and_grid<R, pf::kWarps>(list, s_lists, &s_is_last, ws_lists, ws_counter + query, query, cta, nctas, warp, lane))
        return;
    finish_query<R>(list, k, query, lane, out_score, out_idx, idx_offset, xchg);
    if (lane == 0) {
        ws_counter[query] = 0u;
        if (out_stats) out_stats[query] = atomicAdd(ws_stats + query, 0u);
    }
}

// --------------------------------------------------------------------------------------- host
static inline int pf_r_for_k(int k) { return k <= 32 ? 1 : k <= 64 ? 2 : k <= 128 ? 4 : 8; }

static int pf_grid_y(int64_t rows_per_cta_iter, int64_t n, int sm_count) {
    int64_t want = static_cast<int64_t>(sm_count) * pf::kCtasPerSm;
    int64_t need = (n + rows_per_cta_iter - 1) / rows_per_cta_iter;
    if (need < 1) need = 1;
    return static_cast<int>(want < need ? want : need);
}

static int64_t pf_u_offset(int nq, int k, int sm_count) {
    const int64_t L = 32 * pf_r_for_k(k);
    const int64_t lists = static_cast<int64_t>(nq) * sm_count * pf::kCtasPerSm * L * 8;
    return ((pf::kOffLists + lists + 255) / 256) * 256;
}

int64_t prefilter_workspace_bytes(int64_t n, int nq, int k, int sm_count) {
    return pf_u_offset(nq, k, sm_count) + static_cast<int64_t>(nq) * (n > 0 ? n : 1) * 4;
}

int launch_quantize_rows(const void* D, int dtype, int64_t n, void* D8, void* meta, int sm_count,
                         cudaStream_t stream) {
    if (n == 0) return 0;
    int64_t blocks = (n + pf::kWarps - 1) / pf::kWarps;
    int64_t grid = static_cast<int64_t>(sm_count) * 8;
    if (grid > blocks) grid = blocks;
    dim3 g(static_cast<unsigned>(grid)), b(pf::kWarps * 32);
    int8_t* d8 = static_cast<int8_t*>(D8);
    float4* mt = static_cast<float4*>(meta);
    switch (dtype) {
        case 0: quantize_rows_kernel<float><<<g, b, 0, stream>>>(static_cast<const float*>(D), n, d8, mt); break;
        case 1: quantize_rows_kernel<__nv_bfloat16><<<g, b, 0, stream>>>(static_cast<const __nv_bfloat16*>(D), n, d8, mt); break;
        case 2: quantize_rows_kernel<__half><<<g, b, 0, stream>>>(static_cast<const __half*>(D), n, d8, mt); break;
        case 3: quantize_rows_kernel<Bf16x2><<<g, b, 0, stream>>>(static_cast<const Bf16x2*>(D), n, d8, mt); break;
        default: set_error("quantize_rows: bad dtype %d", dtype); return -1;
    }
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) { set_error("quantize_rows: launch: %s", cudaGetErrorString(e)); return -2; }
    return 0;
}

template <typename T, int R>
static int launch_prefiltered_t(const void* D, int64_t n, const void* D8, const void* meta, const void* Q,
                                bool raw_q, int nq, int k, float* out_score, int64_t* out_idx, int64_t idx_offset,
                                unsigned* out_stats, void* ws, int sm_count, const XchgArgs& xchg, bool pdl,
                                cudaStream_t stream) {
    char* w = static_cast<char*>(ws);
    unsigned* counters = reinterpret_cast<unsigned*>(w + pf::kOffCounters);
    float* tau = reinterpret_cast<float*>(w + pf::kOffTau);
    unsigned* stats = reinterpret_cast<unsigned*>(w + pf::kOffStats);
    uint64_t* lists = reinterpret_cast<uint64_t*>(w + pf::kOffLists);
    float* U = reinterpret_cast<float*>(w + pf_u_offset(nq, k, sm_count));
    cudaLaunchConfig_t cfg = {};
    cfg.blockDim = dim3(pf::kWarps * 32);
    cfg.stream = stream;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = attr;
    const int8_t* d8 = static_cast<const int8_t*>(D8);
    const float4* mt = static_cast<const float4*>(meta);
    const T* Dp = static_cast<const T*>(D);
    const float* Uc = U;
    const float* tauc = tau;
    cudaError_t e;
    // pass A may start while the previous kernel of the stream is in its tail only when the caller
    // vouches for the queries (raw-query form); pass B may always be SCHEDULED early: it waits for
    // pass A at its top, which hides its launch latency
    cfg.gridDim = dim3(nq, pf_grid_y(pf::kWarps * pf::kRowsPerIter, n, sm_count));
    cfg.numAttrs = (pdl && raw_q) ? 1 : 0;
    if (raw_q)
        e = cudaLaunchKernelEx(&cfg, coarse_scan_kernel<T, R, true>, d8, mt, n, Q, k, lists, counters, tau, stats, U);
    else
        e = cudaLaunchKernelEx(&cfg, coarse_scan_kernel<T, R, false>, d8, mt, n, Q, k, lists, counters, tau, stats, U);
    if (e != cudaSuccess) { set_error("prefiltered: scan launch: %s", cudaGetErrorString(e)); return -2; }
    cfg.gridDim = dim3(nq, pf_grid_y(pf::kWarps * 32 * 8, n, sm_count));
    cfg.numAttrs = 1;
    if (raw_q)
        e = cudaLaunchKernelEx(&cfg, rescore_kernel<T, R, true>, Dp, n, Q, k, Uc, tauc, stats, lists, counters,
                               out_score, out_idx, idx_offset, out_stats, xchg);
    else
        e = cudaLaunchKernelEx(&cfg, rescore_kernel<T, R, false>, Dp, n, Q, k, Uc, tauc, stats, lists, counters,
                               out_score, out_idx, idx_offset, out_stats, xchg);
    if (e != cudaSuccess) { set_error("prefiltered: rescore launch: %s", cudaGetErrorString(e)); return -2; }
    return 0;
}

template <typename T>
static int launch_prefiltered_r(const void* D, int64_t n, const void* D8, const void* meta, const void* Q,
                                bool raw_q, int nq, int k, float* out_score, int64_t* out_idx, int64_t idx_offset,
                                unsigned* out_stats, void* ws, int sm_count, const XchgArgs& xchg, bool pdl,
                                cudaStream_t stream) {
    switch (pf_r_for_k(k)) {
        case 1: return launch_prefiltered_t<T, 1>(D, n, D8, meta, Q, raw_q, nq, k, out_score, out_idx, idx_offset, out_stats, ws, sm_count, xchg, pdl, stream);
        case 2: return launch_prefiltered_t<T, 2>(D, n, D8, meta, Q, raw_q, nq, k, out_score, out_idx, idx_offset, out_stats, ws, sm_count, xchg, pdl, stream);
        case 4: return launch_prefiltered_t<T, 4>(D, n, D8, meta, Q, raw_q, nq, k, out_score, out_idx, idx_offset, out_stats, ws, sm_count, xchg, pdl, stream);
        default: return launch_prefiltered_t<T, 8>(D, n, D8, meta, Q, raw_q, nq, k, out_score, out_idx, idx_offset, out_stats, ws, sm_count, xchg, pdl, stream);
    }
}

int launch_topk_prefiltered(const void* D, int dtype, int64_t n, const void* D8, const void* meta,
                            const void* Q, bool raw_q, int nq, int k, float* out_score, int64_t* out_idx,
                            int64_t idx_offset, unsigned* out_stats, void* ws, int64_t ws_bytes,
                            int sm_count, cudaStream_t stream, const XchgArgs* xchg_in, bool pdl) {
    XchgArgs none;
    memset(&none, 0, sizeof(none));
    const XchgArgs& xchg = xchg_in ? *xchg_in : none;
    if (nq > pf::kMaxQueries) { set_error("prefiltered: at most %d queries per call", pf::kMaxQueries); return -1; }
    if (ws_bytes < prefilter_workspace_bytes(n, nq, k, sm_count)) {
        set_error("prefiltered: workspace %lld < %lld bytes", (long long)ws_bytes,
                  (long long)prefilter_workspace_bytes(n, nq, k, sm_count));
        return -3;
    }
    switch (dtype) {
        case 0: return launch_prefiltered_r<float>(D, n, D8, meta, Q, raw_q, nq, k, out_score, out_idx, idx_offset, out_stats, ws, sm_count, xchg, pdl, stream);
        case 1: return launch_prefiltered_r<__nv_bfloat16>(D, n, D8, meta, Q, raw_q, nq, k, out_score, out_idx, idx_offset, out_stats, ws, sm_count, xchg, pdl, stream);
        case 2: return launch_prefiltered_r<__half>(D, n, D8, meta, Q, raw_q, nq, k, out_score, out_idx, idx_offset, out_stats, ws, sm_count, xchg, pdl, stream);
        case 3: return launch_prefiltered_r<Bf16x2>(D, n, D8, meta, Q, raw_q, nq, k, out_score, out_idx, idx_offset, out_stats, ws, sm_count, xchg, pdl, stream);
        default: set_error("prefiltered: bad dtype %d", dtype); return -1;
    }
}

}  // namespace sqe
