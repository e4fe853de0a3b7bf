// K3  batch-1 exact cosine top-k: an HBM-bound streaming GEMV with a warp-resident
// top-k, plus K4 (merge of per-shard lists) and the K5 threshold epilogue.
//
// Replaces the k-NN request of OpenSearchIndexer.search (app/main.py:356-367, external
// HNSW) by exact scoring of every stored row, and -- with k = 1 -- the scan of
// lfu_cache_get (app/main.py:73-90).
//
// Layout: the shard is [n, 1024] row-major in HBM (2 KB rows for bf16/fp16, 4 KB for
// fp32).  One warp scores one row at a time: every lane issues 128-bit
// `ld.global.nc.L1::no_allocate` loads that are contiguous across the warp (512 B per
// instruction), RPW rows are in flight per warp (16 outstanding 128-bit loads per lane),
// two 256-thread CTAs per SM => ~128 KB in flight per SM, enough to cover HBM latency at
// 6.5 TB/s.  The query sits in 32 fp32 registers per lane in the same element layout.
// Scores are fp32 FMA chains (two accumulators per lane, then a 5-step butterfly), so a
// row's score does not depend on where the row sits -- duplicate rows tie exactly and
// the composite key resolves the tie to the lower row.
//
// Selection: each warp keeps a sorted WarpList (sqe_common.cuh); a row is inserted only
// if its key beats the list's worst entry (rare after the first few hundred rows).  The
// eight warp lists of a CTA are merged by bitonic merges, each CTA publishes one list,
// and the last CTA to finish (atomic ticket) merges all CTA lists and writes the result
// -- one launch per query batch, no second kernel.
//
// Algorithmic bytes per query: n * 1024 * sizeof(T)  (+ 1024*sizeof(T) query, + 12*k out).
#include "sqe_common.cuh"
#include "sqe_internal.h"
#include "sqe_rowload.cuh"
#include "sqe_select.cuh"

#include <cstring>

namespace sqe {

constexpr int kGemvWarps = 8;
constexpr int kGemvCtasPerSm = 2;

// RAWQ: `Qv` holds the RAW fp32 query embeddings; every CTA normalises its query itself
// (x / (|x| + 1e-9), the K1 arithmetic bit for bit, app/main.py:353-354) and rounds it to the
// shard's storage type -- the separate K1 launch for one query disappears.
template <typename T, int R, bool RAWQ>
__global__ void __launch_bounds__(kGemvWarps * 32, kGemvCtasPerSm)
topk_gemv_kernel(const T* __restrict__ D, int64_t n, const void* __restrict__ Qv, int k,
                 uint64_t* __restrict__ ws_lists, unsigned* __restrict__ ws_counter,
                 float* __restrict__ out_score, int64_t* __restrict__ out_idx,
                 int64_t idx_offset, const XchgArgs xchg) {
    using E = Elem<T>;
    constexpr int LOADS = E::kLoads;
    constexpr int PER = E::kPer;
    constexpr int QG = E::kQGroups;
    constexpr int RPW = 16 / LOADS;              // rows in flight per warp
    constexpr int L = 32 * R;
    __shared__ uint64_t s_lists[kGemvWarps][L];
    __shared__ int s_is_last;

    const int lane = threadIdx.x & 31;
    const int warp = threadIdx.x >> 5;
    // grid = (queries, CTAs per query): the query is the FAST index, so the CTAs resident at any
    // moment are the same row ranges for all queries -- a row leaves HBM once and the other
    // queries' CTAs read it from L2 (one HBM pass for a small batch instead of one per query)
    const int query = blockIdx.x;
    const int cta = blockIdx.y;
    const int nctas = gridDim.y;

    // the query in registers, same element layout as a row's loads
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

    WarpList<R> list;
    list.clear();
    uint64_t worst = 0ull;

    const int64_t warps_total = static_cast<int64_t>(nctas) * kGemvWarps;
    const int64_t gw = static_cast<int64_t>(cta) * kGemvWarps + warp;
    for (int64_t base = gw * RPW; base < n; base += warps_total * RPW) {
        uint4 raw[RPW][LOADS];
#pragma unroll
        for (int j = 0; j < RPW; ++j) {
            const int64_t row = base + j;
            if (row < n) {
                const T* rp = D + row * E::kRowElems + lane * PER;
#pragma unroll
                for (int c = 0; c < LOADS; ++c) raw[j][c] = ldg_stream(rp + E::load_off(c));
            } else {
#pragma unroll
                for (int c = 0; c < LOADS; ++c) raw[j][c] = make_uint4(0, 0, 0, 0);
            }
        }
        float s[RPW];
#pragma unroll
        for (int j = 0; j < RPW; ++j) {
            float a0 = 0.f, a1 = 0.f;
#pragma unroll
            for (int g = 0; g < QG; ++g) {
                float f[PER];
                elem_group<T>(raw[j], g, f);
#pragma unroll
                for (int e = 0; e < PER; e += 2) {
                    a0 = fmaf(f[e], q[g * PER + e], a0);
                    a1 = fmaf(f[e + 1], q[g * PER + e + 1], a1);
                }
            }
            s[j] = a0 + a1;
        }
#pragma unroll
        for (int d = 16; d >= 1; d >>= 1) {
#pragma unroll
            for (int j = 0; j < RPW; ++j) s[j] += __shfl_xor_sync(kFull, s[j], d);
        }
#pragma unroll
        for (int j = 0; j < RPW; ++j) {
            const int64_t row = base + j;
            if (row < n) {
                const uint64_t key = make_key(s[j], static_cast<uint32_t>(row));
                if (key > worst) {                   // warp-uniform
                    list.insert(key, lane);
                    worst = list.worst();
                }
            }
        }
    }

    // the scan is over: the next scan of the stream may start, and everything below touches
    // memory shared with earlier kernels (sqe_select.cuh)
    pdl_launch_dependents();
    pdl_wait();

    // ---- 8 warp lists -> CTA list -> the query's last CTA merges all CTA lists (sqe_select.cuh);
    // in the sharded mode that same warp also exchanges the result with the other ranks ----
    if (!merge_cta_and_grid<R, kGemvWarps>(list, s_lists, &s_is_last, ws_lists, ws_counter + query, query, cta,
                                           nctas, warp, lane))
        return;
    finish_query<R>(list, k, query, lane, out_score, out_idx, idx_offset, xchg);
    if (lane == 0) ws_counter[query] = 0u;
}

static inline int r_for_k(int k) { return k <= 32 ? 1 : k <= 64 ? 2 : k <= 128 ? 4 : 8; }

static int gemv_grid_x(int64_t n, int sm_count, int rpw) {
    int64_t want = static_cast<int64_t>(sm_count) * kGemvCtasPerSm;
    int64_t need = (n + static_cast<int64_t>(kGemvWarps) * rpw - 1) / (static_cast<int64_t>(kGemvWarps) * rpw);
    if (need < 1) need = 1;
    return static_cast<int>(want < need ? want : need);
}

// workspace = [counters: one u32 per query, padded to 4 KB][per-CTA lists].  The counters of
// up to 1024 queries always sit in the first 4 KB; the kernel leaves them at zero, so they
// need zeroing only once (header contract) -- no memset launch per call.
static int64_t gemv_counter_bytes(int nq) { return ((static_cast<int64_t>(nq) * 4 + 4095) / 4096) * 4096; }

int64_t gemv_workspace_bytes(int nq, int k, int sm_count) {
    const int64_t L = 32 * r_for_k(k);
    const int64_t lists = static_cast<int64_t>(nq) * sm_count * kGemvCtasPerSm * L * 8;
    return gemv_counter_bytes(nq) + lists;
}

template <typename T, int R>
static int launch_gemv_t(const void* D, int64_t n, const void* Q, bool raw_q, int nq, int k,
                         float* out_score, int64_t* out_idx, int64_t idx_offset, void* ws, int sm_count,
                         const XchgArgs& xchg, bool pdl, cudaStream_t stream) {
    constexpr int RPW = 16 / Elem<T>::kLoads;
    const int gx = gemv_grid_x(n, sm_count, RPW);
    unsigned* ws_counter = static_cast<unsigned*>(ws);
    uint64_t* ws_lists = reinterpret_cast<uint64_t*>(static_cast<char*>(ws) + gemv_counter_bytes(nq));
    cudaError_t e;
    if (nq > 1024) {          // counters beyond the reserved 4 KB may hold an older call's lists
        e = cudaMemsetAsync(ws_counter, 0, static_cast<size_t>(nq) * 4, stream);
        if (e != cudaSuccess) { set_error("gemv: memset: %s", cudaGetErrorString(e)); return -2; }
    }
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(nq, gx);
    cfg.blockDim = dim3(kGemvWarps * 32);
    cfg.stream = stream;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = attr;
    cfg.numAttrs = (pdl && raw_q) ? 1 : 0;          // only the raw-query form may start early
    const T* Dp = static_cast<const T*>(D);
    if (raw_q)
        e = cudaLaunchKernelEx(&cfg, topk_gemv_kernel<T, R, true>, Dp, n, Q, k, ws_lists, ws_counter, out_score,
                               out_idx, idx_offset, xchg);
    else
        e = cudaLaunchKernelEx(&cfg, topk_gemv_kernel<T, R, false>, Dp, n, Q, k, ws_lists, ws_counter, out_score,
                               out_idx, idx_offset, xchg);
    if (e != cudaSuccess) { set_error("gemv: launch: %s", cudaGetErrorString(e)); return -2; }
    return 0;
}

template <typename T>
static int launch_gemv_r(const void* D, int64_t n, const void* Q, bool raw_q, int nq, int k, float* out_score,
                         int64_t* out_idx, int64_t idx_offset, void* ws, int sm_count,
                         const XchgArgs& xchg, bool pdl, cudaStream_t stream) {
    switch (r_for_k(k)) {
        case 1: return launch_gemv_t<T, 1>(D, n, Q, raw_q, nq, k, out_score, out_idx, idx_offset, ws, sm_count, xchg, pdl, stream);
        case 2: return launch_gemv_t<T, 2>(D, n, Q, raw_q, nq, k, out_score, out_idx, idx_offset, ws, sm_count, xchg, pdl, stream);
        case 4: return launch_gemv_t<T, 4>(D, n, Q, raw_q, nq, k, out_score, out_idx, idx_offset, ws, sm_count, xchg, pdl, stream);
        default: return launch_gemv_t<T, 8>(D, n, Q, raw_q, nq, k, out_score, out_idx, idx_offset, ws, sm_count, xchg, pdl, stream);
    }
}

int make_xchg_args(XchgArgs* x, int rank, int world, void* const* peer_buffers, int64_t cap, unsigned epoch,
                   int nq, int k) {
    memset(x, 0, sizeof(*x));
    if (world <= 1 || peer_buffers == nullptr) return 0;
    if (world > kMaxWorld || rank < 0 || rank >= world) { set_error("exchange: bad rank %d / world %d", rank, world); return -1; }
    if (nq > kXchgFusedQueries) { set_error("fused exchange: at most %d queries per call (got %d)", kXchgFusedQueries, nq); return -1; }
    if (cap < static_cast<int64_t>(nq) * k) { set_error("exchange: capacity %lld < %d entries", (long long)cap, nq * k); return -1; }
    for (int g = 0; g < world; ++g) {
        if (!peer_buffers[g]) { set_error("exchange: peer buffer %d is null", g); return -1; }
        x->peers.p[g] = static_cast<char*>(peer_buffers[g]);
    }
    x->cap = cap;
    x->rank = rank;
    x->world = world;
    x->epoch = epoch;
    return 0;
}

int launch_topk_gemv(const void* D, int dtype, int64_t n, const void* Q, bool raw_q, int nq, int k,
                     float* out_score, int64_t* out_idx, int64_t idx_offset, void* ws,
                     int64_t ws_bytes, int sm_count, cudaStream_t stream, const XchgArgs* xchg, bool pdl) {
    if (ws_bytes < gemv_workspace_bytes(nq, k, sm_count)) {
        set_error("gemv: workspace %lld < %lld bytes", (long long)ws_bytes,
                  (long long)gemv_workspace_bytes(nq, k, sm_count));
        return -3;
    }
    XchgArgs none;
    memset(&none, 0, sizeof(none));
    const XchgArgs& x = xchg ? *xchg : none;
    switch (dtype) {
        case 0: return launch_gemv_r<float>(D, n, Q, raw_q, nq, k, out_score, out_idx, idx_offset, ws, sm_count, x, pdl, stream);
        case 1: return launch_gemv_r<__nv_bfloat16>(D, n, Q, raw_q, nq, k, out_score, out_idx, idx_offset, ws, sm_count, x, pdl, stream);
        case 2: return launch_gemv_r<__half>(D, n, Q, raw_q, nq, k, out_score, out_idx, idx_offset, ws, sm_count, x, pdl, stream);
        case 3: return launch_gemv_r<Bf16x2>(D, n, Q, raw_q, nq, k, out_score, out_idx, idx_offset, ws, sm_count, x, pdl, stream);
        default: set_error("gemv: bad dtype %d", dtype); return -1;
    }
}

// ---------------------------------------------------------------------------
// K4  merge of per-shard lists: one warp per query.
// ---------------------------------------------------------------------------
template <int R>
__global__ void __launch_bounds__(128)
merge_topk_kernel(const float* __restrict__ scores, const int64_t* __restrict__ idx, int lists,
                  int b, int k_in, int k_out, float* __restrict__ out_score,
                  int64_t* __restrict__ out_idx) {
    const int lane = threadIdx.x & 31;
    const int query = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if (query >= b) return;
    WarpList<R> list;
    list.clear();
    for (int l = 0; l < lists; ++l) {
        const int64_t base = (static_cast<int64_t>(l) * b + query) * k_in;
        WarpList<R> other;
#pragma unroll
        for (int r = 0; r < R; ++r) {
            const int i = r * 32 + lane;
            uint64_t key = 0ull;
            if (i < k_in) {
                const int64_t gi = idx[base + i];
                if (gi >= 0) key = make_key(scores[base + i], static_cast<uint32_t>(gi));
            }
            other.key[r] = key;
        }
        // a well-formed input list is already sorted; sorting again makes the merge
        // independent of that assumption (cheap: lists * log^2 L compare-exchanges)
        other.sort(lane);
        list.merge_sorted(other.key, lane);
    }
    emit_topk<R>(list, k_out, lane, out_score + static_cast<int64_t>(query) * k_out,
                 out_idx + static_cast<int64_t>(query) * k_out, 0);
}

int launch_merge_topk(const float* scores, const int64_t* idx, int lists, int b, int k_in,
                      int k_out, float* out_score, int64_t* out_idx, cudaStream_t stream) {
    if (b == 0) return 0;
    const int kmax = k_in > k_out ? k_in : k_out;
    dim3 block(128), grid((b + 3) / 4);
    switch (r_for_k(kmax)) {
        case 1: merge_topk_kernel<1><<<grid, block, 0, stream>>>(scores, idx, lists, b, k_in, k_out, out_score, out_idx); break;
        case 2: merge_topk_kernel<2><<<grid, block, 0, stream>>>(scores, idx, lists, b, k_in, k_out, out_score, out_idx); break;
        case 4: merge_topk_kernel<4><<<grid, block, 0, stream>>>(scores, idx, lists, b, k_in, k_out, out_score, out_idx); break;
        default: merge_topk_kernel<8><<<grid, block, 0, stream>>>(scores, idx, lists, b, k_in, k_out, out_score, out_idx); break;
    }
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) { set_error("merge: launch: %s", cudaGetErrorString(e)); return -2; }
    return 0;
}

// ---------------------------------------------------------------------------
// K5 epilogue: threshold test of lfu_cache_get (app/main.py:74-75,84,89-90).
// ---------------------------------------------------------------------------
__global__ void cache_finalize_kernel(const float* __restrict__ score, const int64_t* __restrict__ idx,
                                      int b, double threshold, float* __restrict__ out_score,
                                      int32_t* __restrict__ out_idx, uint8_t* __restrict__ out_hit) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= b) return;
    const float s = score[i];
    const int64_t r = idx[i];
    // the reference's running maximum starts at (-1.0, index -1) and only a strictly
    // larger similarity replaces it
    const bool valid = (r >= 0) && (s > -1.0f);
    out_score[i] = valid ? s : -1.0f;
    out_idx[i] = valid ? static_cast<int32_t>(r) : -1;
    out_hit[i] = (valid && !(static_cast<double>(s) < threshold)) ? 1 : 0;   // Python-float compare
}

int launch_cache_finalize(const float* score, const int64_t* idx, int b, double threshold,
                          float* out_score, int32_t* out_idx, uint8_t* out_hit,
                          cudaStream_t stream) {
    if (b == 0) return 0;
    cache_finalize_kernel<<<(b + 127) / 128, 128, 0, stream>>>(score, idx, b, threshold, out_score,
                                                               out_idx, out_hit);
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) { set_error("cache_finalize: launch: %s", cudaGetErrorString(e)); return -2; }
    return 0;
}

}  // namespace sqe
