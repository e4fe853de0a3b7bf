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
//   d-tile  = 256 consecutive shard rows (UMMA N = 256 = 256 fp32 TMEM columns)
//   q-tile  = 128 queries per CTA (the 128 TMEM lanes: lane i = query i); in the CTA-pair
//             form (cta_group::2, UMMA M = 256) a q-tile is 256 queries, 128 per CTA
//   A CTA (pair) owns ONE q-tile for its whole life and walks d-tiles group, group+G, ...
//   (G = #SMs / CTAs per q-tile / #q-tiles).  The units of one group work on the same d-tile
//   at the same time, so a d-tile leaves HBM once and is re-read from L2 by the other q-tiles
//   (ncu: dram bytes = 1.006 x shard bytes).
//
// Warp roles (224 threads, one CTA per SM)
//   warp 0   TMA producer: per K-chunk of 64 elements (= one 128-byte swizzle row) loads this
//            CTA's Q chunk [128 x 64] and its (share of the) D chunk [256 or 128 x 64] into an
//            smem ring; in pair mode both CTAs' bytes are counted on the leader's mbarrier
//   warp 1   TMEM allocator + MMA issuer (leader CTA only): one thread issues 4 x tcgen05.mma
//            (K = 16 each) per stage, 64 per tile, into one of two TMEM accumulators
//            (2 x 256 columns = all 512); tcgen05.commit (multicast to both CTAs in pair mode)
//            releases the smem stage / hands the accumulator to the epilogue
//   warps 2-5 epilogue: thread = one query (TMEM lane), tcgen05.ld 32 columns at a time.
//            Fast path: max tree of the 32 scores against the query's threshold, one vote.
//            Slow path: passing scores are appended (branch-free) as 64-bit keys to a small
//            per-query smem buffer; pending buffers are folded into the query's sorted top
//            list at the end of the tile by the whole warp (bitonic sort sized to the candidate
//            count + bitonic merge, or plain insertion for one or two candidates).  Lists of up
//            to 32 keys live in shared memory and are written through to the workspace.
//            First tile: a register-only bootstrap pass derives a lower bound per query so the
//            empty lists are not rebuilt 256 times.
//   warp 6   threshold warp: keeps merging the partial lists of ALL groups for its share of the
//            q-tile's queries and publishes the k-th best key's score (atomicMax) -- the exact
//            k-th best over everything merged so far on the whole GPU.  Every epilogue thread
//            filters with max(own list's k-th best, that shared bound).
//   A second small kernel merges the G partial lists of every query and writes (score, row).
//
// Exactness: a score below a valid lower bound of the final k-th best can never be in the
// result; ties are resolved by the composite key (score desc, row asc) in every sort/merge
// (sqe_common.cuh).  The local filter is strict (a CTA visits rows in increasing order, so
// an equal score always has a larger row than what the list holds); the shared bound is
// applied non-strictly (thr = just below it) because another CTA's rows may be smaller.
//
// Roofline: tensor pipe (b > ~200) or HBM (small b).  Algorithmic FLOPs per launch =
// 2 * b * n * 1024; algorithmic bytes = n * 2048.  Measured (10M x 1024 bf16, b = 1024, one
// B200 at its 1000 W cap): 1130-1250 TFLOP/s, the same as the main loop without any epilogue
// and 94-97 % of a cuBLAS GEMM of that shape that does no selection.
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

// CG = CTAs cooperating on one UMMA (tcgen05 cta_group).  CG = 1: M = 128, the CTA loads the
// whole 256-row D chunk.  CG = 2: a CTA pair (cluster of 2) computes M = 256; each CTA loads
// its 128 query rows and HALF of the D chunk (128 rows), the tensor cores of both SMs read
// both halves -- 2/3 of the smem and L2 traffic per FLOP of the CG = 1 form.
//
// R = 64-bit keys per lane of a query's sorted list (list length 32 R >= k).  For R = 1
// (k <= 32) the 128 lists of the CTA live in shared memory (32 KB, one operand stage less)
// and are written through to the global workspace; longer lists live in the workspace only.
//
// TOP1 = the k = 1 specialisation (cache lookup): no lists, no candidate buffers, one more
// operand stage (this case is HBM-bound: more bytes in flight).
//
// DEEP (pair form only) = the deep operand ring.  When several q-tiles share d-tiles the ring is
// kept SHALLOW (4 stages) on purpose: the unit that touches a d-tile first pays the DRAM
// latency, and with only 4 stages that slows it down just enough for the followers (L2 hits) to
// stay bunched behind it.  With 5+ stages the leaders never slow down, the units drift apart
// (16 us after 200 tiles), tiles fall out of L2 before the laggards read them and 1.6x the shard
// comes from DRAM -- measured 10M x 1024, b = 1024: 5 stages 32.5 GB / 55.6 k q/s, 4 stages
// 21.5 GB / 60.2 k q/s, 3 stages 21.6 GB but a starved tensor pipe (73 % active).  A pair that
// has its d-tiles to itself (one q-tile in flight) keeps the deep ring.
//
// NARROW (single-CTA form, k <= 16): the smem-resident lists keep only their best 16 keys, which
// frees a fourth operand stage -- this form serves small batches and is HBM-bound, so bytes in
// flight are what counts (3 stages: 5.7 TB/s, 4 stages: 6.5 TB/s).
template <int CG, int R = 1, bool TOP1 = false, bool DEEP = false, bool NARROW = false>
struct Cfg {
    static constexpr int kQTile = kRowsPerCta * CG;          // queries per q-tile
    static constexpr int kBRows = kTileN / CG;               // D rows this CTA loads per chunk
    static constexpr int kBBytes = kBRows * kChunkK * 2;
    static constexpr int kStageBytes = kABytes + kBBytes;    // 48 KB / 32 KB
    static constexpr bool kSmemLists = (R == 1) && !TOP1;
    static constexpr int kListWidth = kSmemLists ? (NARROW ? 16 : 32) : 0;     // keys per list in smem
    static constexpr int kListBytes = kRowsPerCta * kListWidth * 8;
    static constexpr int kBufBytes = TOP1 ? 0 : kRowsPerCta * kBufStride * 8;
    static constexpr int kStages = (CG == 1) ? ((kSmemLists && !NARROW) ? 3 : 4)
                                   : !DEEP    ? 4
                                              : (kSmemLists ? 5 : (TOP1 ? 7 : 6));
    static constexpr int kOffLists = kStages * kStageBytes;
    static constexpr int kOffBuf = kOffLists + kListBytes;
    static constexpr int kOffBar = kOffBuf + kBufBytes;
    static constexpr int kOffTmemPtr = kOffBar + 24 * 8;       // u32 tmem base, u32 epilogue-done counter
    static constexpr int kSmemBytes = kOffTmemPtr + 16 + 1024;   // + alignment slack
};
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
    // diagnostics (sqe_debug_k2_timers)
    unsigned n_slow;      // strips that took the slow path
    unsigned n_flush;     // query lists merged by this warp
    unsigned n_cols;      // column branches taken in the slow path
    long long t_flush;    // cycles inside flush_lanes
};

// pass iff score > thr: strictly above the local k-th best, at or above the shared bound.
// "At or above g" is "above the next smaller float"; the code just below +0.0 is -0.0, which
// compares EQUAL to +0.0, so step once more (to the largest negative denormal).
__device__ __forceinline__ float thr_of(float tau_l, uint32_t tau_g) {
    uint32_t o = tau_g - 1u;
    if (o == 0x7fffffffu) o = 0x7ffffffeu;
    const float g = tau_g ? from_orderable_u32(o) : __int_as_float(0xff800000);
    return fmaxf(tau_l, g);
}

// Merge the pending candidates of every lane in `mask` into that query's sorted list.
// SL = the master copy of the list is in shared memory (`slists`, this warp's 32 x 32 keys)
// and the global copy is write-only here; otherwise the list of the next pending query is
// fetched from L2 while the current one is sorted/merged.
template <int R, int LW>
__device__ __forceinline__ void flush_lanes(unsigned mask, EpiState& st, const uint64_t* wbuf,
                                            uint64_t* slists, uint64_t* wlists, uint32_t* wtau,
                                            int k, int lane) {
    constexpr int L = 32 * R;
    constexpr bool SL = LW > 0;          // LW = keys per list kept in smem (0: lists only in the workspace)
    const long long t_in = clock64();
    st.n_flush += __popc(mask);
    __syncwarp();                                               // owners' buffer stores are visible
    int r = __ffs(mask) - 1;
    mask &= mask - 1;
    uint64_t nxt[R];
    if constexpr (!SL) {
#pragma unroll
        for (int i = 0; i < R; ++i) nxt[i] = __ldcg(wlists + static_cast<size_t>(r) * L + i * 32 + lane);
    }
    while (true) {
        WarpList<R> cur;
        const int r_next = mask ? (__ffs(mask) - 1) : -1;
        mask &= mask - 1;
        if constexpr (SL) {
            cur.key[0] = (lane < LW) ? slists[r * LW + lane] : 0ull;
        } else {
#pragma unroll
            for (int i = 0; i < R; ++i) cur.key[i] = nxt[i];
            if (r_next >= 0) {
#pragma unroll
                for (int i = 0; i < R; ++i)
                    nxt[i] = __ldcg(wlists + static_cast<size_t>(r_next) * L + i * 32 + lane);
            }
        }
        const int c = __shfl_sync(kFull, st.cnt, r);
        if (c <= 2) {
            // the common case once the thresholds are tight: one or two candidates -> plain
            // sorted insertion (the list's worst entry is checked first)
            const uint64_t c0 = wbuf[r * k2::kBufStride];
            const uint64_t c1 = (c == 2) ? wbuf[r * k2::kBufStride + 1] : 0ull;
            if (c0 > cur.worst()) cur.insert(c0, lane);
            if (c1 > cur.worst()) cur.insert(c1, lane);
        } else {
            WarpList<1> cand;
            cand.key[0] = (lane < c) ? wbuf[r * k2::kBufStride + lane] : 0ull;
            if (c <= 4) cand.sort_prefix<4>(lane);
            else if (c <= 8) cand.sort_prefix<8>(lane);
            else if (c <= 16) cand.sort_prefix<16>(lane);
            else cand.sort(lane);
            uint64_t other[R];
#pragma unroll
            for (int i = 0; i < R; ++i) other[i] = (i == 0) ? cand.key[0] : 0ull;
            cur.merge_sorted(other, lane);
        }
        if constexpr (SL) {
            if (lane < LW) slists[r * LW + lane] = cur.key[0];
        }
        uint64_t* lp = wlists + static_cast<size_t>(r) * L;
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
                atomicMax(wtau + r, o);                          // result unused: a RED, no round trip
                st.tau_g = max(st.tau_g, o);
            }
            st.thr = thr_of(st.tau_l, st.tau_g);
        }
        if (r_next < 0) break;
        r = r_next;
    }
    __syncwarp();
    st.t_flush += clock64() - t_in;
}

__device__ __forceinline__ float fmax3(float a, float b, float c) { return fmaxf(fmaxf(a, b), c); }

// One 32-column strip of the accumulator: v[j] = score of (this thread's query, row col0+j).
template <int R, int LW>
__device__ __forceinline__ void process_strip(uint32_t (&v)[32], uint32_t col0, uint32_t n,
                                              bool row_valid, EpiState& st, uint64_t* wbuf,
                                              uint64_t* slists, uint64_t* wlists, uint32_t* wtau,
                                              int k, int lane) {
    if (col0 + 32u > n) {                                       // ragged last d-tile (warp-uniform)
#pragma unroll
        for (int j = 0; j < 32; ++j)
            if (col0 + j >= n) v[j] = 0xff800000u;              // -inf never passes
    }
    float f[32];
#pragma unroll
    for (int j = 0; j < 32; ++j) f[j] = __uint_as_float(v[j]);
    // fast path: a shallow max tree (depth 4) and one vote
    float t[11];
#pragma unroll
    for (int i = 0; i < 10; ++i) t[i] = fmax3(f[3 * i], f[3 * i + 1], f[3 * i + 2]);
    t[10] = fmaxf(f[30], f[31]);
    const float u0 = fmax3(t[0], t[1], t[2]), u1 = fmax3(t[3], t[4], t[5]);
    const float u2 = fmax3(t[6], t[7], t[8]), u3 = fmaxf(t[9], t[10]);
    const float m = fmaxf(fmaxf(u0, u1), fmaxf(u2, u3));
    const bool want = row_valid && (m > st.thr);
    if (!__any_sync(kFull, want)) return;

    // slow path, one half-strip (16 columns = one candidate buffer) at a time: make room where
    // needed (merge the pending candidates of the lanes that would overflow), then branch-free
    // predicated appends.  Works at any pass rate; normally neither half needs a merge.
    ++st.n_slow;
    uint64_t* mybuf = wbuf + lane * k2::kBufStride;
#pragma unroll
    for (int h = 0; h < 2; ++h) {
        const float thr0 = st.thr;
        int pc = 0;
#pragma unroll
        for (int j = 16 * h; j < 16 * h + 16; ++j) pc += (want && f[j] > thr0) ? 1 : 0;
        if (!__any_sync(kFull, pc > 0)) continue;
        const unsigned over = __ballot_sync(kFull, st.cnt + pc > k2::kCap);
        if (over) flush_lanes<R, LW>(over, st, wbuf, slists, wlists, wtau, k, lane);   // their cnt > 0
#pragma unroll
        for (int j = 16 * h; j < 16 * h + 16; ++j) {
            if (want && f[j] > thr0) {
                mybuf[st.cnt] = make_key(f[j], col0 + j);
                ++st.cnt;
            }
        }
    }
}

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

// Bootstrap for k > 16 (first d-tile of a CTA): the first 32*R columns become the query's
// list as they are (unsorted), then the warp sorts each of its 32 lists once.  Without this
// every one of those columns goes through the candidate buffers and k/16 full merges.
template <int R, bool SL>
__device__ __forceinline__ void direct_fill_strip(const uint32_t (&v)[32], uint32_t col0, uint32_t n,
                                                  bool row_valid, int c, uint64_t* slists,
                                                  uint64_t* dst, int lane) {
    constexpr int L = 32 * R;
#pragma unroll
    for (int j = 0; j < 32; ++j) {
        const uint64_t key = (row_valid && col0 + j < n) ? make_key(__uint_as_float(v[j]), col0 + j) : 0ull;
        if constexpr (SL) slists[lane * 32 + j] = key;
        else __stcg(dst + static_cast<size_t>(lane) * L + (c % R) * 32 + j, key);
    }
}

// `src` = where the unsorted batch was written (the list itself for the first batch, the
// scratch list afterwards); `merge` = fold the sorted batch into the existing list.
template <int R, bool SL>
__device__ __forceinline__ void sort_filled_lists(EpiState& st, uint64_t* slists, uint64_t* wlists,
                                                  const uint64_t* src, bool merge, uint32_t* wtau,
                                                  int k, int lane) {
    constexpr int L = 32 * R;
    __syncwarp();                                               // the owners' stores are visible
    uint64_t nxt[R];
    if constexpr (!SL) {
#pragma unroll
        for (int i = 0; i < R; ++i) nxt[i] = __ldcg(src + i * 32 + lane);
    }
#pragma unroll 1
    for (int r = 0; r < 32; ++r) {
        WarpList<R> cur;
        if constexpr (SL) {
            cur.key[0] = slists[r * 32 + lane];
        } else {
#pragma unroll
            for (int i = 0; i < R; ++i) cur.key[i] = nxt[i];
            if (r + 1 < 32) {
#pragma unroll
                for (int i = 0; i < R; ++i)
                    nxt[i] = __ldcg(src + static_cast<size_t>(r + 1) * L + i * 32 + lane);
            }
        }
        cur.sort(lane);
        if constexpr (!SL) {
            if (merge) {
                uint64_t old[R];
#pragma unroll
                for (int i = 0; i < R; ++i) old[i] = __ldcg(wlists + static_cast<size_t>(r) * L + i * 32 + lane);
                cur.merge_sorted(old, lane);
            }
        }
        if constexpr (SL) slists[r * 32 + lane] = cur.key[0];
#pragma unroll
        for (int i = 0; i < R; ++i) __stcg(wlists + static_cast<size_t>(r) * L + i * 32 + lane, cur.key[i]);
        uint64_t kth_src = 0ull;
#pragma unroll
        for (int i = 0; i < R; ++i)
            if (i == ((k - 1) >> 5)) kth_src = cur.key[i];
        const uint64_t kth = shfl_u64(kth_src, (k - 1) & 31);
        if (lane == r && kth != 0ull) {
            st.tau_l = key_score(kth);
            const uint32_t o = static_cast<uint32_t>(kth >> 32);
            atomicMax(wtau + r, o);
            st.tau_g = max(st.tau_g, o);
            st.thr = thr_of(st.tau_l, st.tau_g);
        }
    }
    __syncwarp();
}

// k = 1 (the cache lookup, K5): no lists at all -- every thread keeps the running maximum of
// its query in two registers.  Strict '>' keeps the first maximum: columns only grow while a
// CTA walks its d-tiles, and inside a strip the lowest column holding the maximum is taken
// (the reference's "first maximum wins", app/main.py:84).
__device__ __forceinline__ void top1_strip(uint32_t (&v)[32], uint32_t col0, uint32_t n,
                                           bool row_valid, float& best, uint32_t& best_col) {
    if (col0 + 32u > n) {
#pragma unroll
        for (int j = 0; j < 32; ++j)
            if (col0 + j >= n) v[j] = 0xff800000u;
    }
    float f[32];
#pragma unroll
    for (int j = 0; j < 32; ++j) f[j] = __uint_as_float(v[j]);
    float t[11];
#pragma unroll
    for (int i = 0; i < 10; ++i) t[i] = fmax3(f[3 * i], f[3 * i + 1], f[3 * i + 2]);
    t[10] = fmaxf(f[30], f[31]);
    const float u0 = fmax3(t[0], t[1], t[2]), u1 = fmax3(t[3], t[4], t[5]);
    const float u2 = fmax3(t[6], t[7], t[8]), u3 = fmaxf(t[9], t[10]);
    const float m = fmaxf(fmaxf(u0, u1), fmaxf(u2, u3));
    if (row_valid && m > best) {                            // rare after the first few tiles
        uint32_t j = 31u;
#pragma unroll
        for (int jj = 30; jj >= 0; --jj)
            if (f[jj] == m) j = jj;
        best = m;
        best_col = col0 + j;
    }
}

// Threshold warp: for its share of this q-tile's queries, merge the partial lists of ALL
// groups and publish the k-th best key's score -- the exact k-th best over every row any
// CTA has merged so far.  Lists are read while their owners rewrite them; every slot is an
// 8-byte atomic store of a key that only ever grows, so a torn list is element-wise below a
// real one and the bound stays valid.
template <int R>
__device__ __forceinline__ void threshold_warp(const uint64_t* ws_lists, uint32_t* ws_tau, int b,
                                               int b_pad, int k, int n_groups, int q_row0,
                                               int rows_in_qtile, int my_id, int n_ids,
                                               volatile uint32_t* done, int lane) {
    constexpr int L = 32 * R;
    if (n_groups < 2) return;
    unsigned sleep_ns = 500;        // the bound moves fast at the start, hardly at all later
    while (true) {
        for (int rl = my_id; rl < rows_in_qtile; rl += n_ids) {
            const int row = q_row0 + rl;
            if (row >= b) break;
            WarpList<R> acc;
            acc.clear();
            for (int g0 = 0; g0 < n_groups; g0 += 4) {
                uint64_t o[4][R];
#pragma unroll
                for (int u = 0; u < 4; ++u) {
#pragma unroll
                    for (int i = 0; i < R; ++i)
                        o[u][i] = (g0 + u < n_groups)
                            ? __ldcg(ws_lists + (static_cast<size_t>(g0 + u) * b_pad + row) * L + i * 32 + lane)
                            : 0ull;
                }
#pragma unroll
                for (int u = 0; u < 4; ++u) acc.merge_sorted(o[u], lane);
            }
            uint64_t kth_src = 0ull;
#pragma unroll
            for (int i = 0; i < R; ++i)
                if (i == ((k - 1) >> 5)) kth_src = acc.key[i];
            const uint64_t kth = shfl_u64(kth_src, (k - 1) & 31);
            if (lane == 0 && kth != 0ull) atomicMax(ws_tau + row, static_cast<uint32_t>(kth >> 32));
        }
        if (*done >= 4u) break;
        __nanosleep(sleep_ns);
        if (sleep_ns < 64000u) sleep_ns *= 2;
    }
}

template <int R, int CG, bool TOP1, bool DEEP, bool NARROW>
__global__ void __launch_bounds__(k2::kThreads, 1)
topk_batched_kernel(const __grid_constant__ CUtensorMap tmap_q,
                    const __grid_constant__ CUtensorMap tmap_d, uint32_t n, int b, int k,
                    int n_qt, int n_groups, int n_dtiles, uint32_t idesc,
                    uint64_t* __restrict__ ws_lists, uint32_t* __restrict__ ws_tau,
                    uint32_t* __restrict__ ws_prog, unsigned long long* __restrict__ dbg, int epi_mode,
                    int d_hint, int window, int passes) {
    using namespace k2;
    using C = Cfg<CG, R, TOP1, DEEP, NARROW>;
    constexpr int L = 32 * R;
    constexpr int kStages = C::kStages;
    extern __shared__ uint8_t smem_raw[];
    const uint32_t raw_addr = ptx::smem_u32(smem_raw);
    const uint32_t base = (raw_addr + 1023u) & ~1023u;         // SWIZZLE_128B atoms are 1024-B aligned
    uint8_t* sm = smem_raw + (base - raw_addr);

    const int warp = __shfl_sync(kFull, static_cast<int>(threadIdx.x >> 5), 0);
    const int lane = threadIdx.x & 31;
    const uint32_t rank = (CG == 2) ? ptx::cluster_ctarank() : 0u;   // 0 = leader (issues the MMAs)
    const int unit = blockIdx.x / CG;                                // CTA (CG=1) or CTA pair (CG=2)
    const int q_tile = unit % n_qt;
    const int group = unit / n_qt;
    const int my_tiles = (group < n_dtiles) ? (n_dtiles - group + n_groups - 1) / n_groups : 0;

    const uint32_t bar_full = base + C::kOffBar;               // [kStages]
    const uint32_t bar_empty = bar_full + 8 * kStages;         // [kStages]
    const uint32_t bar_tfull = bar_empty + 8 * kStages;        // [2]
    const uint32_t bar_tempty = bar_tfull + 16;                // [2]
    uint32_t* tmem_ptr_smem = reinterpret_cast<uint32_t*>(sm + C::kOffTmemPtr);
    uint32_t* epi_done = tmem_ptr_smem + 1;

    if (warp == 0 && lane == 0) {
        ptx::prefetch_tensormap(&tmap_q);
        ptx::prefetch_tensormap(&tmap_d);
        for (int s = 0; s < kStages; ++s) {
            ptx::mbar_init(bar_full + 8 * s, 1);               // the leader's expect_tx arrival
            ptx::mbar_init(bar_empty + 8 * s, 1);              // one tcgen05.commit
        }
        for (int a = 0; a < 2; ++a) {
            ptx::mbar_init(bar_tfull + 8 * a, 1);              // one tcgen05.commit
            ptx::mbar_init(bar_tempty + 8 * a, 4 * CG);        // one arrival per epilogue warp (of both CTAs)
        }
        *epi_done = 0u;
        ptx::fence_barrier_init();
    }
    if (warp == 1) ptx::tmem_alloc<CG>(ptx::smem_u32(tmem_ptr_smem), kTmemCols);
    ptx::tc_fence_before();
    if constexpr (CG == 2) ptx::cluster_sync_all(); else __syncthreads();
    ptx::tc_fence_after();
    const uint32_t tmem_base = *tmem_ptr_smem;

    if (warp == 0) {
        // ------------------------------------------------------------ TMA producer
        if (lane == 0) {
            const uint64_t pol_q = ptx::policy_evict_last();   // 2 MB of queries: keep in L2
            // shard rows: no hint by default (an explicit evict_normal policy is classed
            // "evict_normal_demote" by the L2 and doubled the DRAM reads, ncu r1i_k2)
            const uint64_t pol_d = d_hint == 2 ? ptx::policy_evict_first()
                                 : d_hint == 3 ? ptx::policy_evict_last() : ptx::policy_evict_normal();
            const int q_row = q_tile * C::kQTile + static_cast<int>(rank) * kRowsPerCta;
            int stage = 0;
            uint32_t phase = 0;
            long long t_wait = 0;
            const long long t_begin = clock64();
            // Progress window (EXPERIMENT, off unless sqe_tuning_set(SQE_TUNE_K2_WINDOW, w > 0)).
            // In pair mode the units of a group drift apart (start times ~15 us apart after 200
            // tiles); a tile may then have left L2 when the laggards ask for it and is read from
            // HBM again (ncu: 1.2-2x the shard).  With the window every unit publishes the tile
            // it is loading and waits when it is more than `window` tiles ahead of a sibling.
            // Measured: DRAM reads drop to 1.05x, but throughput drops 8-12 % at any window size
            // and the units drift TO the limit instead of staying in their natural ~2-tile
            // equilibrium (followers hit L2 and catch up with the DRAM-fetching leader).
            const uint32_t kWindow = static_cast<uint32_t>(window);
            uint32_t* my_prog = ws_prog + group * n_qt;
            const bool sync_group = (rank == 0) && (n_qt > 1) && (window > 0);
            uint32_t sib[8];
            long long t_window = 0;
            unsigned n_blocked = 0;
            for (int i = 0; i < my_tiles; ++i) {
                const int d_row = (group + i * n_groups) * kTileN + static_cast<int>(rank) * C::kBRows;
                if (sync_group) {
                    asm volatile("st.relaxed.gpu.global.u32 [%0], %1;" ::"l"(my_prog + q_tile), "r"(static_cast<uint32_t>(i + 1)) : "memory");
#pragma unroll
                    for (int u = 0; u < 8; ++u)
                        if (u < n_qt)
                            asm volatile("ld.relaxed.gpu.global.u32 %0, [%1];" : "=r"(sib[u]) : "l"(my_prog + u) : "memory");
                }
                // split-bf16 shards: four passes over K into the same accumulator, smallest
                // terms first -- (Q plane, D plane) = (lo, lo), (hi, lo), (lo, hi), (hi, hi).  The
                // tensor core truncates when it aligns a product group to the accumulator, so the
                // error grows with (#updates x |partial sum|): the cross terms are added while
                // the sum is ~2^-9, and the big hi.hi pass costs what a plain bf16 shard costs.
                for (int ch = 0; ch < kNumChunks * passes; ++ch) {
                    const int kc = ch & (kNumChunks - 1);
                    const int pass = ch / kNumChunks + (4 - passes);     // plain shards: only (hi, hi)
                    const int q_plane = (pass == 0 || pass == 2) ? 1 : 0;
                    const int d_plane = (pass == 0 || pass == 1) ? 1 : 0;
                    const long long w0 = dbg ? clock64() : 0;
                    ptx::mbar_wait(bar_empty + 8 * stage, phase ^ 1u);
                    if (dbg) t_wait += clock64() - w0;
                    const uint32_t sa = base + stage * C::kStageBytes;
                    if constexpr (CG == 1) {
                        const uint32_t fb = bar_full + 8 * stage;
                        ptx::mbar_expect_tx(fb, C::kStageBytes);
                        if (d_hint == 4) ptx::tma_load_3d(sa, &tmap_q, kc * kChunkK, q_plane, q_row, fb);
                        else ptx::tma_load_3d_hint(sa, &tmap_q, kc * kChunkK, q_plane, q_row, fb, pol_q);
                        if (d_hint == 0 || d_hint == 4) ptx::tma_load_3d(sa + kABytes, &tmap_d, kc * kChunkK, d_plane, d_row, fb);
                        else ptx::tma_load_3d_hint(sa + kABytes, &tmap_d, kc * kChunkK, d_plane, d_row, fb, pol_d);
                    } else {
                        // both CTAs' bytes are counted on the LEADER's barrier
                        if (rank == 0) ptx::mbar_expect_tx(bar_full + 8 * stage, 2 * C::kStageBytes);
                        const uint32_t fb = ptx::mapa(bar_full + 8 * stage, 0);
                        if (d_hint == 4) ptx::tma_load_3d_cg2_nohint(sa, &tmap_q, kc * kChunkK, q_plane, q_row, fb);
                        else ptx::tma_load_3d_cg2(sa, &tmap_q, kc * kChunkK, q_plane, q_row, fb, pol_q);
                        if (d_hint == 0 || d_hint == 4) ptx::tma_load_3d_cg2_nohint(sa + kABytes, &tmap_d, kc * kChunkK, d_plane, d_row, fb);
                        else ptx::tma_load_3d_cg2(sa + kABytes, &tmap_d, kc * kChunkK, d_plane, d_row, fb, pol_d);
                    }
                    if (++stage == kStages) { stage = 0; phase ^= 1u; }
                }
                if (sync_group && i + 1 < my_tiles) {
                    // may tile i+1 start?  every sibling must have reached tile i+1-kWindow
                    // (published value = tile index + 1; a finished unit publishes 0xffffffff)
                    const long long t0 = clock64();
                    while (true) {
                        uint32_t lo = 0xffffffffu;
#pragma unroll
                        for (int u = 0; u < 8; ++u)
                            if (u < n_qt) lo = min(lo, sib[u]);
                        if (lo + kWindow >= static_cast<uint32_t>(i + 2) || lo == 0xffffffffu) break;
                        ++n_blocked;
#pragma unroll
                        for (int u = 0; u < 8; ++u)
                            if (u < n_qt)
                                asm volatile("ld.relaxed.gpu.global.u32 %0, [%1];" : "=r"(sib[u]) : "l"(my_prog + u) : "memory");
                        if (clock64() - t0 > 4000000000LL) __trap();
                    }
                    t_window += clock64() - t0;
                }
            }
            if (sync_group)       // done: never hold the others back
                asm volatile("st.relaxed.gpu.global.u32 [%0], %1;" ::"l"(my_prog + q_tile), "r"(0xffffffffu) : "memory");
            if (dbg) {
                dbg[blockIdx.x * 40 + 0] = clock64() - t_begin;
                dbg[blockIdx.x * 40 + 1] = t_wait;
                dbg[blockIdx.x * 40 + 32] = t_window;
                dbg[blockIdx.x * 40 + 33] = n_blocked;
            }
        }
    } else if (warp == 1) {
        // -------------------------------------------------------------- MMA issuer
        if (lane == 0 && rank == 0) {
            int stage = 0;
            uint32_t phase = 0;
            long long t_wfull = 0, t_wtempty = 0;
            const long long t_begin = clock64();
            for (int i = 0; i < my_tiles; ++i) {
                const int acc = i & 1;
                const uint32_t acc_phase = (i >> 1) & 1;
                const long long w0 = dbg ? clock64() : 0;
                if (dbg && (i == 8 || i == 64 || i == 200)) {            // skew between the units of a group
                    unsigned long long gt;
                    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(gt));
                    dbg[blockIdx.x * 40 + (i == 8 ? 5 : i == 64 ? 6 : 7)] = gt;
                }
                ptx::mbar_wait(bar_tempty + 8 * acc, acc_phase ^ 1u);   // epilogue(s) drained it
                if (dbg) t_wtempty += clock64() - w0;
                ptx::tc_fence_after();
                const uint32_t tmem_d = tmem_base + acc * kTileN;
                for (int kc = 0; kc < kNumChunks * passes; ++kc) {
                    const long long w1 = dbg ? clock64() : 0;
                    ptx::mbar_wait(bar_full + 8 * stage, phase);        // TMA bytes have landed
                    if (dbg) t_wfull += clock64() - w1;
                    ptx::tc_fence_after();
                    const uint32_t sa = base + stage * C::kStageBytes;
                    const uint64_t da = make_sw128_desc(sa);
                    const uint64_t db = make_sw128_desc(sa + kABytes);
#pragma unroll
                    for (int k4 = 0; k4 < kChunkK / kUmmaK; ++k4) {
                        // advance 16 elements = 32 bytes inside the swizzle row: +2 in 16-B units
                        ptx::umma_f16<CG>(tmem_d, da + 2 * k4, db + 2 * k4, idesc,
                                          (kc | k4) != 0 ? 1u : 0u);
                    }
                    // frees the smem stage (in both CTAs) once these MMAs have read it
                    if constexpr (CG == 1) ptx::umma_commit(bar_empty + 8 * stage);
                    else ptx::umma_commit_cg2(bar_empty + 8 * stage, 0x3);
                    if (++stage == kStages) { stage = 0; phase ^= 1u; }
                }
                // accumulator complete: wake the epilogue warps (of both CTAs)
                if constexpr (CG == 1) ptx::umma_commit(bar_tfull + 8 * acc);
                else ptx::umma_commit_cg2(bar_tfull + 8 * acc, 0x3);
            }
            if (dbg) {
                dbg[blockIdx.x * 40 + 2] = clock64() - t_begin;
                dbg[blockIdx.x * 40 + 3] = t_wfull;
                dbg[blockIdx.x * 40 + 4] = t_wtempty;
            }
        }
    } else if (warp == 6) {
        // ---------------------------------------------------------- threshold warp
        if constexpr (!TOP1)                                 // k = 1 keeps no lists
        threshold_warp<R>(ws_lists, ws_tau, b, n_qt * C::kQTile, k, n_groups, q_tile * C::kQTile,
                          C::kQTile, group * CG + static_cast<int>(rank), n_groups * CG, epi_done, lane);
    } else {
        // ---------------------------------------------------------------- epilogue
        const int quarter = warp & 3;                                    // TMEM lanes 32q..32q+31
        const int row0 = q_tile * C::kQTile + static_cast<int>(rank) * kRowsPerCta + quarter * 32;
        const bool row_valid = row0 + lane < b;
        const int b_pad = n_qt * C::kQTile;
        constexpr bool SL = C::kSmemLists;
        constexpr int LW = C::kListWidth;
        uint64_t* wbuf = reinterpret_cast<uint64_t*>(sm + C::kOffBuf) + quarter * 32 * kBufStride;
        uint64_t* slists = reinterpret_cast<uint64_t*>(sm + C::kOffLists) + quarter * 32 * LW;
        if constexpr (SL) {
            for (int i = lane; i < 32 * LW; i += 32) slists[i] = 0ull;
            __syncwarp();
        }
        uint64_t* wlists = ws_lists + (static_cast<size_t>(group) * b_pad + row0) * L;
        // scratch lists of the bootstrap (R > 1 only), behind the n_groups * b_pad real lists
        uint64_t* wscratch = wlists + static_cast<size_t>(n_groups) * b_pad * L;
        uint32_t* wtau = ws_tau + row0;

        EpiState st;
        st.tau_l = __int_as_float(0xff800000);
        st.tau_g = 0u;
        st.thr = st.tau_l;
        st.cnt = 0;
        st.n_slow = st.n_flush = st.n_cols = 0u;
        st.t_flush = 0;
        long long t_wtfull = 0, t_ld = 0;
        const long long t_begin = clock64();

        if constexpr (TOP1) {
            float best = __int_as_float(0xff800000);
            uint32_t best_col = 0u;
            for (int i = 0; i < my_tiles; ++i) {
                const int t = group + i * n_groups;
                const int acc = i & 1;
                ptx::mbar_wait(bar_tfull + 8 * acc, (i >> 1) & 1);
                ptx::tc_fence_after();
                const uint32_t taddr = tmem_base + (static_cast<uint32_t>(quarter * 32) << 16) + acc * kTileN;
#pragma unroll 1
                for (int c = 0; c < kTileN / 32; ++c) {
                    uint32_t v[32];
                    ptx::tmem_ld_32x32(taddr + c * 32, v);
                    ptx::tmem_wait_ld();
                    top1_strip(v, static_cast<uint32_t>(t) * kTileN + c * 32, n, row_valid, best, best_col);
                }
                ptx::tc_fence_before();
                __syncwarp();
                if (lane == 0) {
                    if constexpr (CG == 1) ptx::mbar_arrive(bar_tempty + 8 * acc);
                    else ptx::mbar_arrive_cluster(bar_tempty + 8 * acc, 0);
                }
            }
            // slot 0 of this query's (otherwise empty) list; the merge kernel does the rest
            if (best > __int_as_float(0xff800000))
                __stcg(wlists + static_cast<size_t>(lane) * L, make_key(best, best_col));
        } else {
        for (int i = 0; i < my_tiles; ++i) {
            const int t = group + i * n_groups;
            const int acc = i & 1;
            const uint32_t acc_phase = (i >> 1) & 1;
            // refresh the shared bound while waiting for the accumulator
            const uint32_t g = __ldcg(wtau + lane);
            const long long w0 = dbg ? clock64() : 0;
            ptx::mbar_wait(bar_tfull + 8 * acc, acc_phase);
            const long long w1 = dbg ? clock64() : 0;
            if (dbg) t_wtfull += w1 - w0;
            ptx::tc_fence_after();
            if (g > st.tau_g) {
                st.tau_g = g;
                st.thr = thr_of(st.tau_l, st.tau_g);
            }
            const uint32_t taddr = tmem_base + (static_cast<uint32_t>(quarter * 32) << 16) + acc * kTileN;
            if (i == 0 && k <= 16 && epi_mode == 0) {
                float top[16];
#pragma unroll
                for (int j = 0; j < 16; ++j) top[j] = __int_as_float(0xff800000);
#pragma unroll 1
                for (int c = 0; c < kTileN / 32; ++c) {
                    uint32_t v[32];
                    ptx::tmem_ld_32x32(taddr + c * 32, v);
                    ptx::tmem_wait_ld();
                    bootstrap_strip(v, static_cast<uint32_t>(t) * kTileN + c * 32, n, top);
                }
                float kth = top[0];
#pragma unroll
                for (int j = 1; j < 16; ++j)
                    if (j == k - 1) kth = top[j];
                if (row_valid && kth > __int_as_float(0xff800000)) {
                    st.tau_g = max(st.tau_g, orderable_u32(kth));       // applied non-strictly
                    st.thr = thr_of(st.tau_l, st.tau_g);
                }
            }
            uint32_t g_next = __ldcg(wtau + lane);
#pragma unroll 1
            for (int c = 0; c < kTileN / 32; ++c) {
                if (epi_mode == 2) break;                                 // diagnostics: MMA + TMA only
                // the shared bound moves fast while the lists fill up: pick it up twice per tile
                // (the load was issued four strips ago, its latency is hidden)
                if ((c & 3) == 0) {
                    if (g_next > st.tau_g) {
                        st.tau_g = g_next;
                        st.thr = thr_of(st.tau_l, st.tau_g);
                    }
                    g_next = __ldcg(wtau + lane);
                }
                uint32_t v[32];
                const long long l0 = dbg ? clock64() : 0;
                ptx::tmem_ld_32x32(taddr + c * 32, v);
                ptx::tmem_wait_ld();
                if (dbg) t_ld += clock64() - l0;
                if (epi_mode == 1) {                                      // diagnostics: TMEM reads only
                    asm volatile("" ::"r"(v[0]), "r"(v[31]));
                    continue;
                }
                if (!NARROW && i == 0 && k > 16 && (c < R || !SL)) {      // bootstrap for k > 16
                    // lists in smem (k <= 32): the first strip only; lists in the workspace:
                    // the whole tile, in batches of 32*R columns through the scratch list
                    uint64_t* dst = (c < R) ? wlists : wscratch;
                    direct_fill_strip<R, SL>(v, static_cast<uint32_t>(t) * kTileN + c * 32, n, row_valid,
                                             c, slists, dst, lane);
                    if ((c % R) == R - 1)
                        sort_filled_lists<R, SL>(st, slists, wlists, dst, c >= R, wtau, k, lane);
                    continue;
                }
                process_strip<R, LW>(v, static_cast<uint32_t>(t) * kTileN + c * 32, n, row_valid, st,
                                     wbuf, slists, wlists, wtau, k, lane);
            }
            ptx::tc_fence_before();
            __syncwarp();
            if (lane == 0) {
                if constexpr (CG == 1) ptx::mbar_arrive(bar_tempty + 8 * acc);
                else ptx::mbar_arrive_cluster(bar_tempty + 8 * acc, 0);   // the leader's barrier
            }
            // the accumulator is released; now fold this tile's candidates into the lists so
            // the threshold warps see them
            const long long w2 = dbg ? clock64() : 0;
            const unsigned pending = __ballot_sync(kFull, st.cnt > 0);
            if (pending) flush_lanes<R, LW>(pending, st, wbuf, slists, wlists, wtau, k, lane);
            if (dbg && blockIdx.x == 0 && warp == 2 && lane == 0 && i < 64) {
                unsigned long long* tr = dbg + gridDim.x * 40 + i * 4;
                tr[0] = w1 - w0;                 // waited for the accumulator
                tr[1] = w2 - w1;                 // strips (until the accumulator was released)
                tr[2] = clock64() - w2;          // tile-end list merges
                tr[3] = __popc(pending);
            }
        }
        }   // k > 1
        __syncwarp();
        if (lane == 0) atomicAdd(epi_done, 1u);
        if (dbg && lane == 0) {
            unsigned long long* d = dbg + blockIdx.x * 40 + 8 + (warp - 2) * 6;
            d[0] = clock64() - t_begin;
            d[1] = t_wtfull;
            d[2] = st.t_flush;
            d[3] = st.n_slow;
            d[4] = st.n_flush;
            d[5] = t_ld;
        }
    }

    // Teardown.  In pair mode neither CTA may exit (or free TMEM) while the other still reads
    // its shared memory / signals its barriers.
    ptx::tc_fence_before();
    if constexpr (CG == 2) ptx::cluster_sync_all(); else __syncthreads();
    if (warp == 1) {
        ptx::tc_fence_after();
        ptx::tmem_dealloc<CG>(tmem_base, kTmemCols);
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

// Merge the per-group partial lists of each query.  Few groups (large batches): one warp per
// query, four queries per CTA.  Many groups (small batches: up to one group per SM): one CTA
// of eight warps per query -- each warp folds every 8th list (the next one is in flight while
// the current one merges), warp 0 folds the eight partial lists.
template <int R>
__global__ void __launch_bounds__(256)
batched_merge_kernel(const uint64_t* __restrict__ ws_lists, uint32_t* __restrict__ ws_tau, int n_groups,
                     int b, int b_pad, int k, int warps_per_query, float* __restrict__ out_score,
                     int64_t* __restrict__ out_idx, int64_t idx_offset) {
    constexpr int L = 32 * R;
    __shared__ uint64_t s_part[8][L];
    clear_header(ws_tau);
    const int lane = threadIdx.x & 31;
    const int warp = threadIdx.x >> 5;
    const int warps = blockDim.x >> 5;
    const int query = (warps_per_query == 1) ? blockIdx.x * warps + warp : blockIdx.x;
    const int sub = (warps_per_query == 1) ? 0 : warp;          // which share of the groups
    const bool active = query < b;
    WarpList<R> list;
    list.clear();
    if (active) {
        uint64_t nxt[R];
        int g = sub;
        if (g < n_groups) {
#pragma unroll
            for (int r = 0; r < R; ++r)
                nxt[r] = __ldcg(ws_lists + (static_cast<size_t>(g) * b_pad + query) * L + r * 32 + lane);
        }
        while (g < n_groups) {
            WarpList<R> other;
#pragma unroll
            for (int r = 0; r < R; ++r) other.key[r] = nxt[r];
            const int gn = g + warps_per_query;
            if (gn < n_groups) {
#pragma unroll
                for (int r = 0; r < R; ++r)
                    nxt[r] = __ldcg(ws_lists + (static_cast<size_t>(gn) * b_pad + query) * L + r * 32 + lane);
            }
            list.merge_sorted(other.key, lane);
            g = gn;
        }
    }
    if (warps_per_query > 1) {
        list.store(s_part[warp], lane);
        __syncthreads();
        if (warp != 0 || !active) return;
#pragma unroll 1
        for (int w = 1; w < warps_per_query; ++w) {
            WarpList<R> other;
            other.load(s_part[w], lane);
            list.merge_sorted(other.key, lane);
        }
    } else if (!active) {
        return;
    }
    emit_topk<R>(list, k, lane, out_score + static_cast<int64_t>(query) * k,
                 out_idx + static_cast<int64_t>(query) * k, idx_offset);
}

// k = 1: every group wrote at most one key (slot 0 of its list); the result is their maximum.
__global__ void __launch_bounds__(128)
top1_merge_kernel(const uint64_t* __restrict__ ws_lists, uint32_t* __restrict__ ws_tau, int n_groups,
                  int b, int b_pad, float* __restrict__ out_score, int64_t* __restrict__ out_idx,
                  int64_t idx_offset) {
    clear_header(ws_tau);
    const int lane = threadIdx.x & 31;
    const int query = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if (query >= b) return;
    uint64_t best = 0ull;
    for (int g = lane; g < n_groups; g += 32)
        best = umax64(best, __ldcg(ws_lists + (static_cast<size_t>(g) * b_pad + query) * 32));
#pragma unroll
    for (int d = 16; d >= 1; d >>= 1) best = umax64(best, shfl_xor_u64(best, d));
    if (lane == 0) {
        out_score[query] = best ? key_score(best) : __int_as_float(0xff800000);
        out_idx[query] = best ? idx_offset + static_cast<int64_t>(key_row(best)) : -1;
    }
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

// [rows, planes, 1024] 16-bit matrix (planes = 1, or 2 for split bf16: hi | lo), box = 64 elements
// (128 B) x 1 plane x box_rows, 128-B swizzle, out-of-range rows read as zeros.
static int make_tile_map(CUtensorMap* map, const void* ptr, int dtype, uint64_t rows, uint32_t box_rows) {
    EncodeTiledFn enc = encode_tiled_fn();
    if (!enc) { set_error("topk_batched: cuTensorMapEncodeTiled not available from the driver"); return -2; }
    const cuuint64_t planes = (dtype == 3) ? 2 : 1;
    const cuuint64_t dims[3] = {static_cast<cuuint64_t>(kDim), planes, rows};
    const cuuint64_t strides[2] = {static_cast<cuuint64_t>(kDim) * 2, static_cast<cuuint64_t>(kDim) * 2 * planes};
    const cuuint32_t box[3] = {static_cast<cuuint32_t>(k2::kChunkK), 1, box_rows};
    const cuuint32_t estr[3] = {1, 1, 1};
    const CUtensorMapDataType dt = (dtype == 2) ? CU_TENSOR_MAP_DATA_TYPE_FLOAT16 : CU_TENSOR_MAP_DATA_TYPE_BFLOAT16;
    CUresult r = enc(map, dt, 3, const_cast<void*>(ptr), dims, strides, box, estr,
                     CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
                     CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) { set_error("topk_batched: cuTensorMapEncodeTiled failed (%d)", static_cast<int>(r)); return -2; }
    return 0;
}

static inline int r_for_k_batched(int k) { return k <= 32 ? 1 : k <= 64 ? 2 : 4; }

static constexpr int64_t kTauBytes = 4096;        // kQueriesPerLaunch * 4
static constexpr int64_t kProgBytes = 4096;       // one u32 of tile progress per unit (<= #SMs)
static constexpr int64_t kHdrBytes = kTauBytes + kProgBytes;

int64_t batched_workspace_bytes(int64_t /*n*/, int /*b*/, int k, int sm_count) {
    const int R = r_for_k_batched(k);
    const int64_t L = 32 * R;
    // R > 1: as many scratch lists again for the bootstrap of the first d-tile
    return kHdrBytes + static_cast<int64_t>(sm_count) * k2::kRowsPerCta * L * 8 * (R > 1 ? 2 : 1);
}

template <int R, int CG, bool TOP1, bool DEEP, bool NARROW = false>
static int launch_batched_r(const CUtensorMap& tq, const CUtensorMap& td, int64_t n, int b, int k,
                            int n_qt, int n_groups, int n_dtiles, uint32_t idesc, uint64_t* ws_lists,
                            uint32_t* ws_tau, uint32_t* ws_prog, int passes, float* out_score, int64_t* out_idx,
                            int64_t idx_offset, cudaStream_t stream) {
    using C = k2::Cfg<CG, R, TOP1, DEEP, NARROW>;
    // per device and cheap: set on every launch (one process may drive several GPUs)
    cudaError_t e = cudaFuncSetAttribute(topk_batched_kernel<R, CG, TOP1, DEEP, NARROW>,
                                         cudaFuncAttributeMaxDynamicSharedMemorySize, C::kSmemBytes);
    if (e != cudaSuccess) { set_error("topk_batched: smem attribute: %s", cudaGetErrorString(e)); return -2; }
    if (n_dtiles > 0) {
        cudaLaunchConfig_t cfg = {};
        cfg.gridDim = dim3(static_cast<unsigned>(n_groups * n_qt * CG));
        cfg.blockDim = dim3(k2::kThreads);
        cfg.dynamicSmemBytes = C::kSmemBytes;
        cfg.stream = stream;
        cudaLaunchAttribute attr[1];
        attr[0].id = cudaLaunchAttributeClusterDimension;
        attr[0].val.clusterDim.x = CG;
        attr[0].val.clusterDim.y = 1;
        attr[0].val.clusterDim.z = 1;
        cfg.attrs = attr;
        cfg.numAttrs = 1;
        e = cudaLaunchKernelEx(&cfg, topk_batched_kernel<R, CG, TOP1, DEEP, NARROW>, tq, td, static_cast<uint32_t>(n), b, k,
                               n_qt, n_groups, n_dtiles, idesc, ws_lists, ws_tau, ws_prog,
                               reinterpret_cast<unsigned long long*>(g_k2_debug), g_k2_epilogue_mode, g_k2_d_hint, g_k2_window, passes);
        if (e != cudaSuccess) { set_error("topk_batched: launch: %s", cudaGetErrorString(e)); return -2; }
    }
    const int mg = n_dtiles > 0 ? n_groups : 0;
    if constexpr (TOP1) {
        top1_merge_kernel<<<(b + 3) / 4, 128, 0, stream>>>(ws_lists, ws_tau, mg, b, n_qt * C::kQTile, out_score,
                                                           out_idx, idx_offset);
    } else if (mg > 32) {
        batched_merge_kernel<R><<<b, 256, 0, stream>>>(ws_lists, ws_tau, mg, b, n_qt * C::kQTile, k, 8, out_score,
                                                       out_idx, idx_offset);
    } else {
        batched_merge_kernel<R><<<(b + 3) / 4, 128, 0, stream>>>(ws_lists, ws_tau, mg, b, n_qt * C::kQTile, k, 1,
                                                                 out_score, out_idx, idx_offset);
    }
    e = cudaGetLastError();
    if (e != cudaSuccess) { set_error("topk_batched: merge launch: %s", cudaGetErrorString(e)); return -2; }
    return 0;
}

template <int CG>
static int launch_batched_cg(const void* D, int dtype, int64_t n, const void* Q, int b, int k,
                             float* out_score, int64_t* out_idx, int64_t idx_offset, void* ws,
                             int sm_count, cudaStream_t stream) {
    using C = k2::Cfg<CG>;
    const int R = r_for_k_batched(k);
    const int64_t L = 32 * R;
    uint32_t* ws_tau = static_cast<uint32_t*>(ws);
    uint32_t* ws_prog = reinterpret_cast<uint32_t*>(static_cast<char*>(ws) + kTauBytes);
    uint64_t* ws_lists = reinterpret_cast<uint64_t*>(static_cast<char*>(ws) + kHdrBytes);
    const int n_dtiles = static_cast<int>((n + k2::kTileN - 1) / k2::kTileN);
    const uint32_t fmt = (dtype == 2) ? 0u : 1u;               // F16 = 0; BF16 = 1 (also split bf16)
    const int passes = (dtype == 3) ? 4 : 1;                   // split bf16: Ql.Dl + Qh.Dl + Ql.Dh + Qh.Dh
    const int64_t q_row_bytes = (dtype == 3) ? 2 * kDim * 2 : kDim * 2;
    // instruction descriptor: D fp32 [4,6)=1, A fmt [7,10), B fmt [10,13), K-major both,
    // N>>3 [17,23), M>>4 [24,29)  (M = 128 per CTA, 256 for the pair)
    const uint32_t idesc = (1u << 4) | (fmt << 7) | (fmt << 10) |
                           (static_cast<uint32_t>(k2::kTileN >> 3) << 17) |
                           (static_cast<uint32_t>(C::kQTile >> 4) << 24);
    CUtensorMap td;
    if (n > 0) {
        int rc = make_tile_map(&td, D, dtype, static_cast<uint64_t>(n), C::kBRows);
        if (rc != 0) return rc;
    } else {
        memset(&td, 0, sizeof(td));
    }
    const int units = sm_count / CG;                           // CTAs or CTA pairs that fit the chip
    for (int q0 = 0; q0 < b; q0 += k2::kQueriesPerLaunch) {
        const int bc = (b - q0 < k2::kQueriesPerLaunch) ? (b - q0) : k2::kQueriesPerLaunch;
        const int n_qt = (bc + C::kQTile - 1) / C::kQTile;
        int n_groups = units / n_qt;
        if (n_groups < 1) n_groups = 1;
        if (n_dtiles > 0 && n_groups > n_dtiles) n_groups = n_dtiles;
        const int64_t used = kHdrBytes + static_cast<int64_t>(n_groups) * n_qt * C::kQTile * L * 8;
        cudaError_t e = cudaMemsetAsync(ws, 0, static_cast<size_t>(used), stream);
        if (e != cudaSuccess) { set_error("topk_batched: memset: %s", cudaGetErrorString(e)); return -2; }
        const char* qp = static_cast<const char*>(Q) + static_cast<int64_t>(q0) * q_row_bytes;
        CUtensorMap tq;
        int rc = make_tile_map(&tq, qp, dtype, static_cast<uint64_t>(bc), k2::kRowsPerCta);
        if (rc != 0) return rc;
        float* os = out_score + static_cast<int64_t>(q0) * k;
        int64_t* oi = out_idx + static_cast<int64_t>(q0) * k;
        // deep ring only when a pair has its d-tiles to itself (see Cfg)
        const bool deep = (CG == 2) && (n_qt == 1);
#define SQE_K2_LAUNCH(R_, TOP1_)                                                                        \
    (deep ? launch_batched_r<R_, CG, TOP1_, (CG == 2)>(tq, td, n, bc, k, n_qt, n_groups, n_dtiles, idesc, \
                                                       ws_lists, ws_tau, ws_prog, passes, os, oi, idx_offset, stream) \
          : launch_batched_r<R_, CG, TOP1_, false>(tq, td, n, bc, k, n_qt, n_groups, n_dtiles, idesc,     \
                                                   ws_lists, ws_tau, ws_prog, passes, os, oi, idx_offset, stream))
        if (k == 1) rc = SQE_K2_LAUNCH(1, true);
        else if (CG == 1 && k <= 16)         // narrow smem lists + a fourth stage (see Cfg)
            rc = launch_batched_r<1, CG, false, false, (CG == 1)>(tq, td, n, bc, k, n_qt, n_groups, n_dtiles, idesc,
                                                                  ws_lists, ws_tau, ws_prog, passes, os, oi, idx_offset, stream);
        else if (R == 1) rc = SQE_K2_LAUNCH(1, false);
        else if (R == 2) rc = SQE_K2_LAUNCH(2, false);
        else rc = SQE_K2_LAUNCH(4, false);
#undef SQE_K2_LAUNCH
        if (rc != 0) return rc;
    }
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
    // CTA pairs whenever more than one 128-query tile is in flight (sqe_tuning_set overrides)
    int cg = g_k2_cta_group;
    if (cg == 0) cg = (b > k2::kRowsPerCta) ? 2 : 1;
    if (cg == 2)
        return launch_batched_cg<2>(D, dtype, n, Q, b, k, out_score, out_idx, idx_offset, ws, sm_count, stream);
    return launch_batched_cg<1>(D, dtype, n, Q, b, k, out_score, out_idx, idx_offset, ws, sm_count, stream);
}

}  // namespace sqe
