// K2  batched exact cosine top-k on the 5th-gen tensor cores (tcgen05 + TMEM + TMA) with
// the top-k selection fused into the accumulator epilogue: the [b, n] score matrix only
// ever exists as TMEM tiles, never in HBM.
//
// Replaces the k-NN request of OpenSearchIndexer.search (app/main.py:356-367, an external
// approximate HNSW index) by exact scoring  S = Q . D^T  of every stored row, for a batch
// of queries, and -- with k = 1 -- the scan of lfu_cache_get (app/main.py:73-90) for a
// stream of queries (K5).
//
// Work decomposition
//   q-tile  = 128 consecutive queries  (UMMA M = 128 = the 128 TMEM lanes: lane i = query i)
//   d-tile  = 256 consecutive shard rows (UMMA N = 256 = 256 fp32 TMEM columns)
//   A CTA owns ONE q-tile for its whole life and walks d-tiles group, group+G, group+2G...
//   (G = #SMs / #q-tiles groups).  The CTAs of one group work on the same d-tile at the
//   same time, so a d-tile leaves HBM once and is re-read from L2 by the other q-tiles.
//
// Warp roles (192 threads, one CTA per SM)
//   warp 0   TMA producer: per K-chunk of 64 elements (= one 128-byte swizzle row) loads
//            the Q chunk [128 x 64] and the D chunk [256 x 64] into a 4-stage smem ring
//   warp 1   TMEM allocator + MMA issuer: one thread issues 4 x tcgen05.mma
//            (128 x 256 x 16) per stage, 64 per tile, accumulating in one of two TMEM
//            accumulators (2 x 256 columns = all 512), tcgen05.commit releases the stage
//   warps 2-5 epilogue: thread = one query (TMEM lane), tcgen05.ld 32 columns at a time.
//            Fast path: max of the 32 scores against the query's running threshold.
//            Slow path: passing scores are appended as 64-bit keys to a small per-query
//            smem buffer; a full buffer is bitonic-sorted by the whole warp and merged into
//            the query's sorted top list (L2-resident workspace), which raises the
//            threshold.  The k-th best score is also published per query with atomicMax so
//            every CTA filters with the best lower bound any CTA has found.
//   A second small kernel merges the G partial lists of every query and writes (score, row).
//
// Exactness: a score below a valid lower bound of the final k-th best can never be in the
// result; ties are resolved by the composite key (score desc, row asc) in every sort/merge
// (sqe_common.cuh).  The local filter is strict (a CTA visits rows in increasing order, so
// an equal score always has a larger row than what the list holds); the shared bound is
// applied non-strictly (thr = just below it) because another CTA's rows may be larger.
//
// Roofline: tensor pipe.  Algorithmic FLOPs per launch = 2 * b * n * 1024.
#include <cuda.h>
#include <cstring>

#include "sqe_common.cuh"
#include "sqe_internal.h"
#include "sqe_ptx.cuh"

namespace sqe {

namespace k2 {
constexpr int kTileM = 128;
constexpr int kTileN = 256;
constexpr int kChunkK = 64;                    // elements per K chunk = 128 bytes = swizzle span
constexpr int kNumChunks = kDim / kChunkK;     // 16
constexpr int kUmmaK = 16;
constexpr int kStages = 4;
constexpr int kABytes = kTileM * kChunkK * 2;  // 16 KB
constexpr int kBBytes = kTileN * kChunkK * 2;  // 32 KB
constexpr int kStageBytes = kABytes + kBBytes;
constexpr int kCap = 16;                       // pending candidates per query
constexpr int kBufStride = 17;                 // u64 per query row in smem (padded)
constexpr int kThreads = 192;
constexpr int kTmemCols = 512;

constexpr int kOffBuf = kStages * kStageBytes;                  // 196608
constexpr int kOffBar = kOffBuf + kTileM * kBufStride * 8;      // +17408
constexpr int kOffTmemPtr = kOffBar + 16 * 8;
constexpr int kSmemBytes = kOffTmemPtr + 16 + 1024;             // + alignment slack
constexpr int kQueriesPerLaunch = 1024;
}  // namespace k2

// smem matrix descriptor of a K-major, 128-byte-swizzled operand tile whose rows are 128 B
// apart and whose 8-row groups are 1024 B apart (exactly what a TMA box {64, rows} with
// CU_TENSOR_MAP_SWIZZLE_128B writes).  Fields: start>>4 [0,14), LBO>>4 [16,30) (unused for
// swizzled K-major, 1), SBO>>4 [32,46) = 64, version [46,48) = 1, layout [61,64) = 2 (SW128).
__device__ __forceinline__ uint64_t make_sw128_desc(uint32_t smem_addr) {
    const uint32_t lo = ((smem_addr & 0x3ffffu) >> 4) | (1u << 16);
    const uint32_t hi = 64u | (1u << 14) | (2u << 29);
    return (static_cast<uint64_t>(hi) << 32) | lo;
}

struct EpiState {
    float thr;        // pass iff score > thr
    float tau_l;      // k-th best score of this CTA's list for this query (-inf until k rows)
    uint32_t tau_g;   // best published bound (orderable u32), 0 = none
    int cnt;          // pending candidates in the smem buffer
};

__device__ __forceinline__ float thr_of(float tau_l, uint32_t tau_g) {
    const float g = tau_g ? from_orderable_u32(tau_g - 1u) : __int_as_float(0xff800000);
    return fmaxf(tau_l, g);
}

// Merge the pending candidates of every lane in `mask` into that query's sorted list.
template <int R>
__device__ __forceinline__ void flush_lanes(unsigned mask, EpiState& st, uint64_t* wbuf,
                                            uint64_t* wlists, uint32_t* wtau, int k, int lane) {
    constexpr int L = 32 * R;
    __syncwarp();                                               // owners' buffer stores are visible
    while (mask) {
        const int r = __ffs(mask) - 1;
        mask &= mask - 1;
        const int c = __shfl_sync(kFull, st.cnt, r);
        WarpList<1> cand;
        cand.key[0] = (lane < c) ? wbuf[r * k2::kBufStride + lane] : 0ull;
        cand.sort(lane);
        uint64_t* lp = wlists + static_cast<size_t>(r) * L;
        WarpList<R> cur;
        uint64_t other[R];
#pragma unroll
        for (int i = 0; i < R; ++i) {
            cur.key[i] = __ldcg(lp + i * 32 + lane);
            other[i] = (i == 0) ? cand.key[0] : 0ull;
        }
        cur.merge_sorted(other, lane);
#pragma unroll
        for (int i = 0; i < R; ++i) __stcg(lp + i * 32 + lane, cur.key[i]);
        uint64_t kth_src = 0ull;
#pragma unroll
        for (int i = 0; i < R; ++i)
            if (i == ((k - 1) >> 5)) kth_src = cur.key[i];
        const uint64_t kth = shfl_u64(kth_src, (k - 1) & 31);
        if (lane == r) {
            st.cnt = 0;
            if (kth != 0ull) {
                st.tau_l = key_score(kth);
                const uint32_t o = static_cast<uint32_t>(kth >> 32);
                const uint32_t old = atomicMax(wtau + r, o);
                st.tau_g = max(st.tau_g, max(old, o));
            }
            st.thr = thr_of(st.tau_l, st.tau_g);
        }
        __syncwarp();
    }
}

// One 32-column strip of the accumulator: v[j] = score of (this thread's query, row col0+j).
template <int R>
__device__ __forceinline__ void process_strip(uint32_t (&v)[32], uint32_t col0, uint32_t n,
                                              bool row_valid, EpiState& st, uint64_t* wbuf,
                                              uint64_t* wlists, uint32_t* wtau, int k, int lane) {
    if (col0 + 32u > n) {                                       // ragged last d-tile (warp-uniform)
#pragma unroll
        for (int j = 0; j < 32; ++j)
            if (col0 + j >= n) v[j] = 0xff800000u;              // -inf never passes
    }
    float m = __uint_as_float(v[0]);
#pragma unroll
    for (int j = 1; j < 32; ++j) m = fmaxf(m, __uint_as_float(v[j]));
    const bool want = row_valid && (m > st.thr);
    if (__ballot_sync(kFull, want) == 0u) return;

    uint64_t* mybuf = wbuf + lane * k2::kBufStride;
#pragma unroll
    for (int g = 0; g < 4; ++g) {
        const unsigned full = __ballot_sync(kFull, st.cnt > k2::kCap - 8);
        if (full) flush_lanes<R>(full, st, wbuf, wlists, wtau, k, lane);
#pragma unroll
        for (int jj = 0; jj < 8; ++jj) {
            const int j = g * 8 + jj;
            const float s = __uint_as_float(v[j]);
            if (want && s > st.thr) {
                mybuf[st.cnt] = make_key(s, col0 + j);
                ++st.cnt;
            }
        }
    }
}

template <int R>
__global__ void __launch_bounds__(k2::kThreads, 1)
topk_batched_kernel(const __grid_constant__ CUtensorMap tmap_q,
                    const __grid_constant__ CUtensorMap tmap_d, uint32_t n, int b, int k,
                    int n_qt, int n_groups, int n_dtiles, uint32_t idesc,
                    uint64_t* __restrict__ ws_lists, uint32_t* __restrict__ ws_tau) {
    using namespace k2;
    constexpr int L = 32 * R;
    extern __shared__ uint8_t smem_raw[];
    const uint32_t raw_addr = ptx::smem_u32(smem_raw);
    const uint32_t base = (raw_addr + 1023u) & ~1023u;         // SWIZZLE_128B atoms are 1024-B aligned
    uint8_t* sm = smem_raw + (base - raw_addr);

    const int warp = __shfl_sync(kFull, static_cast<int>(threadIdx.x >> 5), 0);
    const int lane = threadIdx.x & 31;
    const int q_tile = blockIdx.x % n_qt;
    const int group = blockIdx.x / n_qt;
    const int my_tiles = (group < n_dtiles) ? (n_dtiles - group + n_groups - 1) / n_groups : 0;

    const uint32_t bar_full = base + kOffBar;                  // [kStages]
    const uint32_t bar_empty = bar_full + 8 * kStages;         // [kStages]
    const uint32_t bar_tfull = bar_empty + 8 * kStages;        // [2]
    const uint32_t bar_tempty = bar_tfull + 16;                // [2]
    uint32_t* tmem_ptr_smem = reinterpret_cast<uint32_t*>(sm + kOffTmemPtr);

    if (warp == 0 && lane == 0) {
        ptx::prefetch_tensormap(&tmap_q);
        ptx::prefetch_tensormap(&tmap_d);
        for (int s = 0; s < kStages; ++s) {
            ptx::mbar_init(bar_full + 8 * s, 1);
            ptx::mbar_init(bar_empty + 8 * s, 1);
        }
        for (int a = 0; a < 2; ++a) {
            ptx::mbar_init(bar_tfull + 8 * a, 1);
            ptx::mbar_init(bar_tempty + 8 * a, 4);             // one arrival per epilogue warp
        }
        ptx::fence_barrier_init();
    }
    if (warp == 1) ptx::tmem_alloc<1>(ptx::smem_u32(tmem_ptr_smem), kTmemCols);
    ptx::tc_fence_before();
    __syncthreads();
    ptx::tc_fence_after();
    const uint32_t tmem_base = *tmem_ptr_smem;

    if (warp == 0) {
        // ------------------------------------------------------------ TMA producer
        if (lane == 0) {
            int stage = 0;
            uint32_t phase = 0;
            for (int i = 0; i < my_tiles; ++i) {
                const int t = group + i * n_groups;
                for (int kc = 0; kc < kNumChunks; ++kc) {
                    ptx::mbar_wait(bar_empty + 8 * stage, phase ^ 1u);
                    const uint32_t fb = bar_full + 8 * stage;
                    const uint32_t sa = base + stage * kStageBytes;
                    ptx::mbar_expect_tx(fb, kStageBytes);
                    ptx::tma_load_2d(sa, &tmap_q, kc * kChunkK, q_tile * kTileM, fb);
                    ptx::tma_load_2d(sa + kABytes, &tmap_d, kc * kChunkK, t * kTileN, fb);
                    if (++stage == kStages) { stage = 0; phase ^= 1u; }
                }
            }
        }
    } else if (warp == 1) {
        // -------------------------------------------------------------- MMA issuer
        if (lane == 0) {
            int stage = 0;
            uint32_t phase = 0;
            for (int i = 0; i < my_tiles; ++i) {
                const int acc = i & 1;
                const uint32_t acc_phase = (i >> 1) & 1;
                ptx::mbar_wait(bar_tempty + 8 * acc, acc_phase ^ 1u);   // epilogue drained it
                ptx::tc_fence_after();
                const uint32_t tmem_d = tmem_base + acc * kTileN;
                for (int kc = 0; kc < kNumChunks; ++kc) {
                    ptx::mbar_wait(bar_full + 8 * stage, phase);        // TMA bytes have landed
                    ptx::tc_fence_after();
                    const uint32_t sa = base + stage * kStageBytes;
                    const uint64_t da = make_sw128_desc(sa);
                    const uint64_t db = make_sw128_desc(sa + kABytes);
#pragma unroll
                    for (int k4 = 0; k4 < kChunkK / kUmmaK; ++k4) {
                        // advance 16 elements = 32 bytes inside the swizzle row: +2 in 16-B units
                        ptx::umma_f16<1>(tmem_d, da + 2 * k4, db + 2 * k4, idesc,
                                         (kc | k4) != 0 ? 1u : 0u);
                    }
                    ptx::umma_commit(bar_empty + 8 * stage);            // frees the smem stage
                    if (++stage == kStages) { stage = 0; phase ^= 1u; }
                }
                ptx::umma_commit(bar_tfull + 8 * acc);                  // accumulator complete
            }
        }
    } else {
        // ---------------------------------------------------------------- epilogue
        const int quarter = warp & 3;                                    // TMEM lanes 32q..32q+31
        const int row_in_tile = quarter * 32 + lane;
        const int row = q_tile * kTileM + row_in_tile;
        const bool row_valid = row < b;
        const int b_pad = n_qt * kTileM;
        uint64_t* wbuf = reinterpret_cast<uint64_t*>(sm + kOffBuf) + quarter * 32 * kBufStride;
        uint64_t* wlists = ws_lists +
            (static_cast<size_t>(group) * b_pad + q_tile * kTileM + quarter * 32) * L;
        uint32_t* wtau = ws_tau + q_tile * kTileM + quarter * 32;

        EpiState st;
        st.tau_l = __int_as_float(0xff800000);
        st.tau_g = 0u;
        st.thr = st.tau_l;
        st.cnt = 0;

        for (int i = 0; i < my_tiles; ++i) {
            const int t = group + i * n_groups;
            const int acc = i & 1;
            const uint32_t acc_phase = (i >> 1) & 1;
            // refresh the shared bound while waiting for the accumulator
            const uint32_t g = __ldcg(wtau + lane);
            ptx::mbar_wait(bar_tfull + 8 * acc, acc_phase);
            ptx::tc_fence_after();
            if (g > st.tau_g) {
                st.tau_g = g;
                st.thr = thr_of(st.tau_l, st.tau_g);
            }
            const uint32_t taddr = tmem_base + (static_cast<uint32_t>(quarter * 32) << 16) + acc * kTileN;
#pragma unroll 1
            for (int c = 0; c < kTileN / 32; ++c) {
                uint32_t v[32];
                ptx::tmem_ld_32x32(taddr + c * 32, v);
                ptx::tmem_wait_ld();
                process_strip<R>(v, static_cast<uint32_t>(t) * kTileN + c * 32, n, row_valid, st,
                                 wbuf, wlists, wtau, k, lane);
            }
            ptx::tc_fence_before();
            __syncwarp();
            if (lane == 0) ptx::mbar_arrive(bar_tempty + 8 * acc);
        }
        const unsigned pending = __ballot_sync(kFull, st.cnt > 0);
        if (pending) flush_lanes<R>(pending, st, wbuf, wlists, wtau, k, lane);
    }

    ptx::tc_fence_before();
    __syncthreads();
    if (warp == 1) {
        ptx::tc_fence_after();
        ptx::tmem_dealloc<1>(tmem_base, kTmemCols);
    }
}

// Merge the per-group partial lists of each query: one warp per query.
template <int R>
__global__ void __launch_bounds__(128)
batched_merge_kernel(const uint64_t* __restrict__ ws_lists, int n_groups, int b, int b_pad, int k,
                     float* __restrict__ out_score, int64_t* __restrict__ out_idx,
                     int64_t idx_offset) {
    constexpr int L = 32 * R;
    const int lane = threadIdx.x & 31;
    const int query = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if (query >= b) return;
    WarpList<R> list;
    list.clear();
    for (int g = 0; g < n_groups; ++g) {
        WarpList<R> other;
        other.load(ws_lists + (static_cast<size_t>(g) * b_pad + query) * L, lane);
        list.merge_sorted(other.key, lane);
    }
    emit_topk<R>(list, k, lane, out_score + static_cast<int64_t>(query) * k,
                 out_idx + static_cast<int64_t>(query) * k, idx_offset);
}

// ------------------------------------------------------------------------------ host
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*,
                                  const cuuint64_t*, const cuuint64_t*, const cuuint32_t*,
                                  const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

// The one driver entry point this library needs, resolved at run time so the .so loads
// (and exports its ABI) on machines without libcuda.
static EncodeTiledFn encode_tiled_fn() {
    static EncodeTiledFn fn = []() -> EncodeTiledFn {
        void* p = nullptr;
        cudaDriverEntryPointQueryResult q = cudaDriverEntryPointSymbolNotFound;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) != cudaSuccess ||
            q != cudaDriverEntryPointSuccess)
            return nullptr;
        return reinterpret_cast<EncodeTiledFn>(p);
    }();
    return fn;
}

// [rows, 1024] 16-bit row-major matrix, box = 64 elements (128 B) x box_rows, 128-B swizzle,
// out-of-range rows read as zeros.
static int make_tile_map(CUtensorMap* map, const void* ptr, int dtype, uint64_t rows, uint32_t box_rows) {
    EncodeTiledFn enc = encode_tiled_fn();
    if (!enc) { set_error("topk_batched: cuTensorMapEncodeTiled not available from the driver"); return -2; }
    const cuuint64_t dims[2] = {static_cast<cuuint64_t>(kDim), rows};
    const cuuint64_t strides[1] = {static_cast<cuuint64_t>(kDim) * 2};
    const cuuint32_t box[2] = {static_cast<cuuint32_t>(k2::kChunkK), box_rows};
    const cuuint32_t estr[2] = {1, 1};
    const CUtensorMapDataType dt = (dtype == 1) ? CU_TENSOR_MAP_DATA_TYPE_BFLOAT16 : CU_TENSOR_MAP_DATA_TYPE_FLOAT16;
    CUresult r = enc(map, dt, 2, const_cast<void*>(ptr), dims, strides, box, estr,
                     CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
                     CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) { set_error("topk_batched: cuTensorMapEncodeTiled failed (%d)", static_cast<int>(r)); return -2; }
    return 0;
}

static inline int r_for_k_batched(int k) { return k <= 32 ? 1 : k <= 64 ? 2 : 4; }

static constexpr int64_t kTauBytes = 4096;        // kQueriesPerLaunch * 4

int64_t batched_workspace_bytes(int64_t /*n*/, int /*b*/, int k, int sm_count) {
    const int64_t L = 32 * r_for_k_batched(k);
    return kTauBytes + static_cast<int64_t>(sm_count) * k2::kTileM * L * 8;
}

template <int R>
static int launch_batched_r(const CUtensorMap& tq, const CUtensorMap& td, int64_t n, int b, int k,
                            int n_qt, int n_groups, int n_dtiles, uint32_t idesc, uint64_t* ws_lists,
                            uint32_t* ws_tau, float* out_score, int64_t* out_idx, int64_t idx_offset,
                            cudaStream_t stream) {
    // per device and cheap: set on every launch (one process may drive several GPUs)
    cudaError_t e = cudaFuncSetAttribute(topk_batched_kernel<R>,
                                         cudaFuncAttributeMaxDynamicSharedMemorySize, k2::kSmemBytes);
    if (e != cudaSuccess) { set_error("topk_batched: smem attribute: %s", cudaGetErrorString(e)); return -2; }
    if (n_dtiles > 0) {
        topk_batched_kernel<R><<<n_groups * n_qt, k2::kThreads, k2::kSmemBytes, stream>>>(
            tq, td, static_cast<uint32_t>(n), b, k, n_qt, n_groups, n_dtiles, idesc, ws_lists, ws_tau);
        e = cudaGetLastError();
        if (e != cudaSuccess) { set_error("topk_batched: launch: %s", cudaGetErrorString(e)); return -2; }
    }
    batched_merge_kernel<R><<<(b + 3) / 4, 128, 0, stream>>>(ws_lists, n_dtiles > 0 ? n_groups : 0, b,
                                                             n_qt * k2::kTileM, k, out_score, out_idx,
                                                             idx_offset);
    e = cudaGetLastError();
    if (e != cudaSuccess) { set_error("topk_batched: merge launch: %s", cudaGetErrorString(e)); return -2; }
    return 0;
}

int launch_topk_batched(const void* D, int dtype, int64_t n, const void* Q, int b, int k,
                        float* out_score, int64_t* out_idx, int64_t idx_offset, void* ws,
                        int64_t ws_bytes, int sm_count, cudaStream_t stream) {
    if (ws_bytes < batched_workspace_bytes(n, b, k, sm_count)) {
        set_error("topk_batched: workspace %lld < %lld bytes", (long long)ws_bytes,
                  (long long)batched_workspace_bytes(n, b, k, sm_count));
        return -3;
    }
    const int R = r_for_k_batched(k);
    const int64_t L = 32 * R;
    uint32_t* ws_tau = static_cast<uint32_t*>(ws);
    uint64_t* ws_lists = reinterpret_cast<uint64_t*>(static_cast<char*>(ws) + kTauBytes);
    const int n_dtiles = static_cast<int>((n + k2::kTileN - 1) / k2::kTileN);
    const uint32_t fmt = (dtype == 1) ? 1u : 0u;               // BF16 = 1, F16 = 0
    // instruction descriptor: D fp32 [4,6)=1, A fmt [7,10), B fmt [10,13), K-major both,
    // N>>3 [17,23), M>>4 [24,29)
    const uint32_t idesc = (1u << 4) | (fmt << 7) | (fmt << 10) |
                           (static_cast<uint32_t>(k2::kTileN >> 3) << 17) |
                           (static_cast<uint32_t>(k2::kTileM >> 4) << 24);
    CUtensorMap td;
    if (n > 0) {
        int rc = make_tile_map(&td, D, dtype, static_cast<uint64_t>(n), k2::kTileN);
        if (rc != 0) return rc;
    } else {
        memset(&td, 0, sizeof(td));
    }
    for (int q0 = 0; q0 < b; q0 += k2::kQueriesPerLaunch) {
        const int bc = (b - q0 < k2::kQueriesPerLaunch) ? (b - q0) : k2::kQueriesPerLaunch;
        const int n_qt = (bc + k2::kTileM - 1) / k2::kTileM;
        int n_groups = sm_count / n_qt;
        if (n_groups < 1) n_groups = 1;
        if (n_dtiles > 0 && n_groups > n_dtiles) n_groups = n_dtiles;
        const int64_t used = kTauBytes + static_cast<int64_t>(n_groups) * n_qt * k2::kTileM * L * 8;
        cudaError_t e = cudaMemsetAsync(ws, 0, static_cast<size_t>(used), stream);
        if (e != cudaSuccess) { set_error("topk_batched: memset: %s", cudaGetErrorString(e)); return -2; }
        const char* qp = static_cast<const char*>(Q) + static_cast<int64_t>(q0) * kDim * 2;
        CUtensorMap tq;
        int rc = make_tile_map(&tq, qp, dtype, static_cast<uint64_t>(bc), k2::kTileM);
        if (rc != 0) return rc;
        float* os = out_score + static_cast<int64_t>(q0) * k;
        int64_t* oi = out_idx + static_cast<int64_t>(q0) * k;
        switch (R) {
            case 1: rc = launch_batched_r<1>(tq, td, n, bc, k, n_qt, n_groups, n_dtiles, idesc, ws_lists, ws_tau, os, oi, idx_offset, stream); break;
            case 2: rc = launch_batched_r<2>(tq, td, n, bc, k, n_qt, n_groups, n_dtiles, idesc, ws_lists, ws_tau, os, oi, idx_offset, stream); break;
            default: rc = launch_batched_r<4>(tq, td, n, bc, k, n_qt, n_groups, n_dtiles, idesc, ws_lists, ws_tau, os, oi, idx_offset, stream); break;
        }
        if (rc != 0) return rc;
    }
    return 0;
}

}  // namespace sqe
