// Pieces of the tensor-core scoring kernels shared by the bf16/fp16 form (topk_batched.cu, K2)
// and the int8 prefilter form (topk_batched_i8.cu, K2p): tile geometry, the smem operand
// descriptor, the threshold rule, the first-tile bootstrap, and the candidate-log machinery
// (append-only per (query, group) logs in the L2-resident workspace, incremental threshold
// warps, cross-group first-tile bound).
#pragma once
#include <cuda.h>
#include <cstring>

#include "sqe_common.cuh"
#include "sqe_internal.h"
#include "sqe_ptx.cuh"

namespace sqe {

namespace k2 {
constexpr int kRowsPerCta = 128;               // queries per CTA = TMEM lanes
constexpr int kTileN = 256;                    // shard rows per d-tile = fp32 TMEM columns per accumulator
constexpr int kChunkK = 64;                    // elements per K chunk = 128 bytes = swizzle span
constexpr int kNumChunks = kDim / kChunkK;     // 16
constexpr int kUmmaK = 16;
constexpr int kABytes = kRowsPerCta * kChunkK * 2;  // 16 KB: this CTA's 128 query rows of one chunk
constexpr int kCap = 16;                       // pending candidates per query
constexpr int kBufStride = 17;                 // u64 per query row in smem (padded)
constexpr int kThreads = 224;                 // TMA, MMA, 4 epilogue warps, threshold warp
constexpr int kTmemCols = 512;
constexpr int kQueriesPerLaunch = 1024;

constexpr int kMaxGroups = 160;                // >= #SMs: groups per q-tile (seen[] words per query)
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

// pass iff score > thr: strictly above the local k-th best, at or above the shared bound.
// "At or above g" is "above the next smaller float"; the code just below +0.0 is -0.0, which
// compares EQUAL to +0.0, so step once more (to the largest negative denormal).
__device__ __forceinline__ float thr_of(float tau_l, uint32_t tau_g) {
    uint32_t o = tau_g - 1u;
    if (o == 0x7fffffffu) o = 0x7ffffffeu;
    const float g = tau_g ? from_orderable_u32(o) : __int_as_float(0xff800000);
    return fmaxf(tau_l, g);
}

__device__ __forceinline__ float fmax3(float a, float b, float c) { return fmaxf(fmaxf(a, b), c); }

// Bootstrap (first d-tile of a CTA, k <= 16): every list is empty, so without help every
// score passes and the lists are rebuilt 256 times.  One extra pass over the accumulator
// keeps, per query (= per lane, in registers, branch-free), the 16 largest of the 64
// maxima of 4 consecutive columns; the k-th largest of those maxima are k distinct scores,
// so its value is a valid lower bound of the final k-th best.  The regular pass that follows
// then lets only ~k scores per query through.
__device__ __forceinline__ void bootstrap_strip(const uint32_t (&v)[32], uint32_t col0, uint32_t n,
                                                float (&top)[16]) {
#pragma unroll
    for (int g = 0; g < 8; ++g) {
        float x = __int_as_float(0xff800000);
#pragma unroll
        for (int e = 0; e < 4; ++e) {
            const float s = (col0 + 4 * g + e < n) ? __uint_as_float(v[4 * g + e]) : __int_as_float(0xff800000);
            x = fmaxf(x, s);
        }
#pragma unroll
        for (int i = 0; i < 16; ++i) {                          // insertion network, descending
            const float hi = fmaxf(top[i], x);
            x = fminf(top[i], x);
            top[i] = hi;
        }
    }
}

// =============================================================================== log mode
// k > 32.  Per (query, group): log[kLogCap] keys + one count word (epoch << 16 | entries;
// entries == 0xffff while the owner compacts).  Writers never fence on the append path: an
// entry is one 8-byte store into a slot that is ZERO until then (memset at launch, re-zeroed by
// a compaction before its count is published), so a reader that sees the count before the entry
// sees a zero and stops there.  Everything a reader can ever see in a slot is the key of a real
// row, so whatever it merges is a set of real, distinct rows and its k-th best is a valid lower
// bound of the final k-th best.
__device__ __forceinline__ uint32_t ld_relaxed_gpu(const uint32_t* p) {
    uint32_t v;
    asm volatile("ld.relaxed.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void st_relaxed_gpu(uint32_t* p, uint32_t v) {
    asm volatile("st.relaxed.gpu.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}

struct LogState {
    float thr;        // pass iff score > thr
    float tau_l;      // k-th best of this thread's own log at its last compaction (-inf before)
    uint32_t tau_g;   // best published bound (orderable u32), 0 = none
    uint32_t cnt;     // entries in this thread's log
    uint32_t epoch;   // compactions so far (16 bits are published)
    unsigned n_slow, n_compact;      // diagnostics
};

// The logs of the lanes in `mask` are full: sort, keep the best k, zero the rest, bump the epoch.
// The k-th best becomes the lane's LOCAL bound (strict: a CTA visits rows in increasing order).
template <int R>
__device__ __forceinline__ void log_compact_lanes(unsigned mask, LogState& st, uint64_t* wlog,
                                                  uint32_t* wcount, int gpad, uint32_t* wtau, int k,
                                                  int lane) {
    constexpr int CAP = 64 * R;
    constexpr int R2 = 2 * R;
    __syncwarp();                                               // the owners' appends are visible
    while (mask) {
        const int r = __ffs(mask) - 1;
        mask &= mask - 1;
        const uint32_t c = __shfl_sync(kFull, st.cnt, r);
        const uint32_t ep = (__shfl_sync(kFull, st.epoch, r) + 1u) & 0xffffu;
        uint64_t* lp = wlog + static_cast<size_t>(r) * CAP;
        uint32_t* cp = wcount + static_cast<size_t>(r) * gpad;
        if (lane == 0) st_relaxed_gpu(cp, (ep << 16) | 0xffffu);          // "compacting": readers skip it
        __threadfence();
        WarpList<R2> w;
#pragma unroll
        for (int i = 0; i < R2; ++i) {
            const uint32_t idx = i * 32 + lane;
            w.key[i] = idx < c ? __ldcg(lp + idx) : 0ull;
        }
        w.sort(lane);
        uint64_t kth_src = 0ull;
#pragma unroll
        for (int i = 0; i < R2; ++i) {
            const int idx = i * 32 + lane;
            __stcg(lp + idx, idx < k ? w.key[i] : 0ull);
            if (i == ((k - 1) >> 5)) kth_src = w.key[i];
        }
        const uint64_t kth = shfl_u64(kth_src, (k - 1) & 31);
        const uint32_t nc = c < static_cast<uint32_t>(k) ? c : static_cast<uint32_t>(k);
        __threadfence();                                        // the rewritten log before its count
        __syncwarp();
        if (lane == 0) st_relaxed_gpu(cp, (ep << 16) | nc);
        if (lane == r) {
            st.cnt = nc;
            st.epoch = ep;
            ++st.n_compact;
            if (kth != 0ull) {
                st.tau_l = key_score(kth);
                const uint32_t o = static_cast<uint32_t>(kth >> 32);
                atomicMax(wtau + r, o);
                st.tau_g = max(st.tau_g, o);
            }
            st.thr = thr_of(st.tau_l, st.tau_g);
        }
    }
    __syncwarp();
}

// First d-tile in log mode: the groups exchange the j-th best of their first 256 rows (`top`, from
// bootstrap_strip: the j-th largest of the maxima of 4 columns = j distinct scores) and every
// query takes the m-th largest of the published values as its first bound: at least m groups hold
// j rows each at or above it, j m >= k.  The wait is bounded and nothing depends on it: a group
// that has not published counts as -inf, fewer than m published values give no bound.
__device__ __forceinline__ void log_boot_exchange(const float (&top)[16], int boot_j, int boot_m,
                                                  bool row_valid, uint32_t* boot_row, int group,
                                                  int n_groups, uint32_t* arrive, LogState& st, int lane) {
    float jth = top[0];
#pragma unroll
    for (int i = 1; i < 16; ++i)
        if (i == boot_j - 1) jth = top[i];
    if (row_valid && jth > __int_as_float(0xff800000)) st_relaxed_gpu(boot_row + group, orderable_u32(jth));
    __threadfence();
    __syncwarp();
    if (lane == 0) {
        atomicAdd(arrive, 1u);
        const long long t0 = clock64();
        while (ld_relaxed_gpu(arrive) < static_cast<uint32_t>(n_groups) && clock64() - t0 < 200000LL) {}
    }
    __syncwarp();
    __threadfence();
    uint32_t best[16];
#pragma unroll
    for (int i = 0; i < 16; ++i) best[i] = 0u;
    for (int g = 0; g < n_groups; g += 4) {                     // rows are padded to a multiple of 32 words
        const uint4 x4 = __ldcg(reinterpret_cast<const uint4*>(boot_row + g));
        const uint32_t xs[4] = {x4.x, x4.y, x4.z, x4.w};
#pragma unroll
        for (int e = 0; e < 4; ++e) {
            uint32_t x = xs[e];
#pragma unroll
            for (int i = 0; i < 16; ++i) {                      // insertion network, descending
                const uint32_t hi = max(best[i], x);
                x = min(best[i], x);
                best[i] = hi;
            }
        }
    }
    uint32_t mth = best[0];
#pragma unroll
    for (int i = 1; i < 16; ++i)
        if (i == boot_m - 1) mth = best[i];
    if (row_valid && mth != 0u) {
        st.tau_g = max(st.tau_g, mth);                          // applied non-strictly (thr_of)
        st.thr = thr_of(st.tau_l, st.tau_g);
    }
}

// Threshold warp, log mode.  For each query it tracks (state in shared memory: the merged best
// 32 R keys + per group how far into that group's log it has read) it polls the count words,
// reads ONLY entries it has not seen, inserts those that beat the list's worst and publishes the
// k-th best.  After a compaction (epoch change) a log is re-read from its start; keys are unique
// per row, so a key already in the list is skipped.
template <int R>
__device__ __forceinline__ void threshold_warp_log(const int CAP, const uint64_t* ws_logs, const uint32_t* ws_counts,
                                                   uint32_t* ws_tau, uint8_t* state, int n_slots, int b,
                                                   int b_pad, int gpad, int k, int n_groups, int q_row0,
                                                   int rows_in_qtile, int my_id, int n_ids,
                                                   volatile uint32_t* done, int lane, int min_groups = 2) {
    constexpr int L = 32 * R;
    constexpr int kSlotBytes = L * 8 + k2::kMaxGroups * 4;
    if (n_groups < min_groups) return;                          // K2 with one group: its own compactions bound it
    int n_own = 0;
    for (int rl = my_id; rl < rows_in_qtile && q_row0 + rl < b && n_own < n_slots; rl += n_ids) ++n_own;
    if (n_own == 0) return;
    for (int i = lane; i < n_own * (kSlotBytes / 4); i += 32) reinterpret_cast<uint32_t*>(state)[i] = 0u;
    __syncwarp();
    unsigned sleep_ns = 500;
    while (true) {
        const bool last = *done >= 4u;                          // one more round once this CTA's epilogue is through
        bool any_new = false;
        for (int t = 0; t < n_own; ++t) {
            const int row = q_row0 + my_id + t * n_ids;
            uint64_t* sl = reinterpret_cast<uint64_t*>(state + static_cast<size_t>(t) * kSlotBytes);
            uint32_t* seen = reinterpret_cast<uint32_t*>(sl + L);
            WarpList<R> acc;
            acc.load(sl, lane);
            // only a key above the current k-th best can move the bound: that is the insertion
            // threshold (`worst`), not the list's last element
            auto kth_of = [&]() {
                uint64_t src = 0ull;
#pragma unroll
                for (int i = 0; i < R; ++i)
                    if (i == ((k - 1) >> 5)) src = acc.key[i];
                return shfl_u64(src, (k - 1) & 31);
            };
            uint64_t worst = kth_of();
            bool changed = false;
            for (int g0 = 0; g0 < n_groups; g0 += 32) {
                // lane l reads the log of group g0 + l: 32 logs are walked in lockstep, four entries
                // per lane in flight, so a round costs a few L2 round trips, not one per group
                const int g = g0 + lane;
                const bool valid = g < n_groups;
                const uint32_t w = valid ? ld_relaxed_gpu(ws_counts + static_cast<size_t>(row) * gpad + g) : 0u;
                const uint32_t sv = valid ? seen[g] : 0u;
                const bool act = valid && w != sv && (w & 0xffffu) != 0xffffu;
                uint32_t pos = (act && (w >> 16) == (sv >> 16)) ? (sv & 0xffffu) : 0u;
                uint32_t end = act ? min(w & 0xffffu, static_cast<uint32_t>(CAP)) : 0u;
                const uint64_t* lp = ws_logs + (static_cast<size_t>(valid ? g : 0) * b_pad + row) * CAP;
                while (__any_sync(kFull, pos < end)) {
                    uint64_t key[4];
#pragma unroll
                    for (int u = 0; u < 4; ++u) key[u] = (pos + u < end) ? __ldcg(lp + pos + u) : 0ull;
#pragma unroll
                    for (int u = 0; u < 4; ++u) {
                        const bool have = pos + u < end;
                        if (have && key[u] == 0ull) end = pos + u;      // not visible yet: stop here this round
                        unsigned pass = __ballot_sync(kFull, pos + u < end && key[u] > worst);
                        while (pass) {
                            const int l = __ffs(pass) - 1;
                            pass &= pass - 1;
                            const uint64_t cand = shfl_u64(key[u], l);
                            if (cand > worst) {
                                bool dup = false;
#pragma unroll
                                for (int i = 0; i < R; ++i) dup |= (acc.key[i] == cand);
                                if (!__any_sync(kFull, dup)) {
                                    acc.insert(cand, lane);
                                    worst = kth_of();
                                    changed = true;
                                }
                            }
                        }
                    }
                    pos = min(pos + 4u, end);
                }
                if (act) seen[g] = (w & 0xffff0000u) | pos;
            }
            if (changed) {
                any_new = true;
                acc.store(sl, lane);
                const uint64_t kth = worst;
                if (lane == 0 && kth != 0ull) atomicMax(ws_tau + row, static_cast<uint32_t>(kth >> 32));
            }
            __syncwarp();
        }
        if (last) break;
        if (any_new) sleep_ns = 500;
        __nanosleep(sleep_ns);
        if (sleep_ns < 8000u) sleep_ns *= 2;
    }
}

// The first 4 KB of a workspace double as the ticket counters of the GEMV kernel when a caller
// shares one workspace between entry points (sqe_cache_top1 does); the header contract
// (include/sqe_b200.h) is that every call leaves them ZERO.  The main kernel has finished
// (stream order) when the merge kernel runs, so the published bounds are dead by now.
__device__ __forceinline__ void clear_header(uint32_t* ws_tau) {
    if (blockIdx.x == 0)
        for (int i = threadIdx.x; i < k2::kQueriesPerLaunch; i += blockDim.x) ws_tau[i] = 0u;
}

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

// [rows, planes, 1024] 16-bit matrix (planes = 1, or 2 for split bf16: hi | lo), box = 64 elements
// (128 B) x 1 plane x box_rows, 128-B swizzle, out-of-range rows read as zeros.
// dtype: the storage codes of include/sqe_b200.h (1 bf16, 2 fp16, 3 split bf16), or kTileMapInt8
// for [rows, 1024] int8 rows (box = 128 elements = the same 128-byte swizzle row).
constexpr int kTileMapInt8 = 100;
static int make_tile_map(CUtensorMap* map, const void* ptr, int dtype, uint64_t rows, uint32_t box_rows) {
    EncodeTiledFn enc = encode_tiled_fn();
    if (!enc) { set_error("topk_batched: cuTensorMapEncodeTiled not available from the driver"); return -2; }
    const bool i8 = dtype == kTileMapInt8;
    const cuuint64_t esize = i8 ? 1 : 2;
    const cuuint64_t planes = (dtype == 3) ? 2 : 1;
    const cuuint64_t dims[3] = {static_cast<cuuint64_t>(kDim), planes, rows};
    const cuuint64_t strides[2] = {static_cast<cuuint64_t>(kDim) * esize, static_cast<cuuint64_t>(kDim) * esize * planes};
    const cuuint32_t box[3] = {static_cast<cuuint32_t>(128 / esize), 1, box_rows};     // 128 bytes = one swizzle row
    const cuuint32_t estr[3] = {1, 1, 1};
    const CUtensorMapDataType dt = i8 ? CU_TENSOR_MAP_DATA_TYPE_UINT8
                                      : (dtype == 2) ? CU_TENSOR_MAP_DATA_TYPE_FLOAT16 : CU_TENSOR_MAP_DATA_TYPE_BFLOAT16;
    CUresult r = enc(map, dt, 3, const_cast<void*>(ptr), dims, strides, box, estr,
                     CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
                     CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) { set_error("topk_batched: cuTensorMapEncodeTiled failed (%d)", static_cast<int>(r)); return -2; }
    return 0;
}

}  // namespace sqe
