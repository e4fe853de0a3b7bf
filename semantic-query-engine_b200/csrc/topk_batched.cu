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
//            k > 32 ("log mode"): no sorted lists at all in the main loop.  A passing score is
//            APPENDED (one 8-byte store, no load, no merge) to the (query, group) candidate log in
//            the L2-resident workspace; the threshold warps consume the logs incrementally and
//            publish the exact k-th best of everything logged so far, so only ~k ln(n/k) scores
//            per query are ever logged.  First tile: the groups exchange the j-th best of their
//            first 256 rows and every query starts from the m-th largest of those (j m >= k).
//   warp 6   threshold warp: keeps merging the partial lists of ALL groups for its share of the
//            q-tile's queries and publishes the k-th best key's score (atomicMax) -- the exact
//            k-th best over everything merged so far on the whole GPU.  Every epilogue thread
//            filters with max(own list's k-th best, that shared bound).  In log mode it keeps
//            the merged top list of each of its queries in shared memory and only reads log
//            entries it has not seen yet.
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
#include "sqe_k2.cuh"

namespace sqe {

namespace k2 {
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
//
// LOG (R > 1, i.e. k > 32) = candidate logs instead of sorted lists (see the file header): per
// (query, group) an append-only array of kLogCap = 64 R keys in the workspace plus one count word
// (epoch << 16 | entries).  When a log is full its owner's warp sorts it, keeps the best k and
// bumps the epoch (rare: the published bounds keep the logs short).  No candidate buffers and no
// lists in shared memory; the space goes to the threshold warp's per-query state instead.
template <int CG, int R = 1, bool TOP1 = false, bool DEEP = false, bool NARROW = false>
struct Cfg {
    static constexpr int kQTile = kRowsPerCta * CG;          // queries per q-tile
    static constexpr int kBRows = kTileN / CG;               // D rows this CTA loads per chunk
    static constexpr int kBBytes = kBRows * kChunkK * 2;
    static constexpr int kStageBytes = kABytes + kBBytes;    // 48 KB / 32 KB
    static constexpr bool kLog = (R > 1) && !TOP1;
    static constexpr int kLogCap = 64 * R;                   // entries per (query, group) log
    static constexpr bool kSmemLists = (R == 1) && !TOP1;
    static constexpr int kListWidth = kSmemLists ? (NARROW ? 16 : 32) : 0;     // keys per list in smem
    static constexpr int kListBytes = kRowsPerCta * kListWidth * 8;
    static constexpr int kBufBytes = (TOP1 || kLog) ? 0 : kRowsPerCta * kBufStride * 8;
    static constexpr int kStages = (CG == 1) ? ((kSmemLists && !NARROW) ? 3 : 4)
                                   : !DEEP    ? 4
                                              : (kSmemLists ? 5 : (TOP1 ? 7 : 6));
    // threshold warp state (log mode): per tracked query a merged list of 32 R keys + one
    // `seen` word per group
    static constexpr int kThrSlotBytes = 32 * R * 8 + kMaxGroups * 4;
    static constexpr int kThrSlots = !kLog ? 0 : ((CG == 1 || DEEP) ? 16 : 48);
    static constexpr int kThrBytes = kThrSlots * kThrSlotBytes;
    static constexpr int kOffLists = kStages * kStageBytes;
    static constexpr int kOffBuf = kOffLists + kListBytes;
    static constexpr int kOffThr = kOffBuf + kBufBytes;
    static constexpr int kOffBar = kOffThr + kThrBytes;
    static constexpr int kOffTmemPtr = kOffBar + 24 * 8;       // u32 tmem base, u32 epilogue-done counter
    static constexpr int kSmemBytes = kOffTmemPtr + 16 + 1024;   // + alignment slack
    static_assert(kSmemBytes <= 227 * 1024, "shared memory budget");
};
}  // namespace k2

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

// Merge the pending candidates of every lane in `mask` into that query's sorted list.
// SL = the master copy of the list is in shared memory (`slists`, this warp's 32 x 32 keys)
// and the global copy is write-only here; otherwise the list of the next pending query is
// fetched from L2 while the current one is sorted/merged.
template <int R, int LW, bool DBG>
__device__ __forceinline__ void flush_lanes(unsigned mask, EpiState& st, const uint64_t* wbuf,
                                            uint64_t* slists, uint64_t* wlists, uint32_t* wtau,
                                            int k, int lane) {
    constexpr int L = 32 * R;
    constexpr bool SL = LW > 0;          // LW = keys per list kept in smem (0: lists only in the workspace)
    long long t_in = 0;
    if constexpr (DBG) {
        t_in = clock64();
        st.n_flush += __popc(mask);
    }
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
    if constexpr (DBG) st.t_flush += clock64() - t_in;
}

// One 32-column strip of the accumulator: v[j] = score of (this thread's query, row col0+j).
template <int R, int LW, bool DBG>
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
    if constexpr (DBG) ++st.n_slow;
    uint64_t* mybuf = wbuf + lane * k2::kBufStride;
#pragma unroll
    for (int h = 0; h < 2; ++h) {
        const float thr0 = st.thr;
        int pc = 0;
#pragma unroll
        for (int j = 16 * h; j < 16 * h + 16; ++j) pc += (want && f[j] > thr0) ? 1 : 0;
        if (!__any_sync(kFull, pc > 0)) continue;
        const unsigned over = __ballot_sync(kFull, st.cnt + pc > k2::kCap);
        if (over) flush_lanes<R, LW, DBG>(over, st, wbuf, slists, wlists, wtau, k, lane);   // their cnt > 0
#pragma unroll
        for (int j = 16 * h; j < 16 * h + 16; ++j) {
            if (want && f[j] > thr0) {
                mybuf[st.cnt] = make_key(f[j], col0 + j);
                ++st.cnt;
            }
        }
    }
}

// Bootstrap for 16 < k <= 32 (first d-tile of a CTA, lists in shared memory): the first 32
// columns become the query's list as they are (unsorted), then the warp sorts each of its 32
// lists once.  Without this every one of those columns goes through the candidate buffers.
__device__ __forceinline__ void direct_fill_strip(const uint32_t (&v)[32], uint32_t col0, uint32_t n,
                                                  bool row_valid, uint64_t* slists, int lane) {
#pragma unroll
    for (int j = 0; j < 32; ++j)
        slists[lane * 32 + j] = (row_valid && col0 + j < n) ? make_key(__uint_as_float(v[j]), col0 + j) : 0ull;
}

__device__ __forceinline__ void sort_filled_lists(EpiState& st, uint64_t* slists, uint64_t* wlists,
                                                  uint32_t* wtau, int k, int lane) {
    __syncwarp();                                               // the owners' stores are visible
#pragma unroll 1
    for (int r = 0; r < 32; ++r) {
        WarpList<1> cur;
        cur.key[0] = slists[r * 32 + lane];
        cur.sort(lane);
        slists[r * 32 + lane] = cur.key[0];
        __stcg(wlists + static_cast<size_t>(r) * 32 + lane, cur.key[0]);
        const uint64_t kth = shfl_u64(cur.key[0], (k - 1) & 31);
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

// One 32-column strip in log mode: same fast path; a passing score is appended to the log.
template <int R>
__device__ __forceinline__ void process_strip_log(uint32_t (&v)[32], uint32_t col0, uint32_t n,
                                                  bool row_valid, LogState& st, uint64_t* wlog,
                                                  uint32_t* wcount, int gpad, uint32_t* wtau, int k,
                                                  int lane) {
    constexpr int CAP = 64 * R;
    if (col0 + 32u > n) {                                       // ragged last d-tile (warp-uniform)
#pragma unroll
        for (int j = 0; j < 32; ++j)
            if (col0 + j >= n) v[j] = 0xff800000u;              // -inf never passes
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
    const bool want = row_valid && (m > st.thr);
    if (!__any_sync(kFull, want)) return;

    ++st.n_slow;
    int pc = 0;
#pragma unroll
    for (int j = 0; j < 32; ++j) pc += (want && f[j] > st.thr) ? 1 : 0;
    const unsigned over = __ballot_sync(kFull, st.cnt + pc > static_cast<uint32_t>(CAP));
    if (over) log_compact_lanes<R>(over, st, wlog, wcount, gpad, wtau, k, lane);   // <= k entries left, thr raised
    if (want) {
        uint64_t* mylog = wlog + static_cast<size_t>(lane) * CAP;
        const float thr0 = st.thr;
        uint32_t c = st.cnt;
#pragma unroll
        for (int j = 0; j < 32; ++j) {
            if (f[j] > thr0) {
                __stcg(mylog + c, make_key(f[j], col0 + j));
                ++c;
            }
        }
        if (c != st.cnt) {
            st.cnt = c;
            st_relaxed_gpu(wcount + static_cast<size_t>(lane) * gpad, (st.epoch << 16) | c);
        }
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

// Launch geometry and workspace pointers of one launch (host -> kernel, by value).
struct K2Args {
    uint32_t n;               // shard rows
    int b, k;
    int n_qt, n_groups, n_dtiles;
    uint32_t idesc;
    int passes;               // 1, or 4 for split-bf16 shards
    uint64_t* ws_lists;       // lists [group][b_pad][32 R]  |  logs [group][b_pad][64 R] (log mode)
    uint32_t* ws_tau;         // published bounds, one per query (first 4 KB of the workspace)
    uint32_t* ws_arrive;      // log mode: bootstrap arrivals per 32-query slice
    uint32_t* ws_counts;      // log mode: [b_pad][gpad] count words
    uint32_t* ws_boot;        // log mode: [b_pad][gpad] j-th best of each group's first d-tile
    int gpad;                 // n_groups rounded up to 32
    int boot_j, boot_m;       // log mode bootstrap (0 = none): m-th largest of the groups' j-th best
    unsigned long long* dbg;  // role timers (DBG instantiations only)
    int epi_mode;             // DBG only: 1 = TMEM loads only, 2 = no epilogue (results invalid)
};

template <int R, int CG, bool TOP1, bool DEEP, bool NARROW, bool DBG>
__global__ void __launch_bounds__(k2::kThreads, 1)
topk_batched_kernel(const __grid_constant__ CUtensorMap tmap_q,
                    const __grid_constant__ CUtensorMap tmap_d, const K2Args a) {
    using namespace k2;
    using C = Cfg<CG, R, TOP1, DEEP, NARROW>;
    constexpr int L = 32 * R;
    constexpr int kStages = C::kStages;
    constexpr bool LOG = C::kLog;
    extern __shared__ uint8_t smem_raw[];
    const uint32_t raw_addr = ptx::smem_u32(smem_raw);
    const uint32_t base = (raw_addr + 1023u) & ~1023u;         // SWIZZLE_128B atoms are 1024-B aligned
    uint8_t* sm = smem_raw + (base - raw_addr);

    const uint32_t n = a.n;
    const int b = a.b, k = a.k, n_qt = a.n_qt, n_groups = a.n_groups, n_dtiles = a.n_dtiles;
    const int passes = a.passes;
    [[maybe_unused]] unsigned long long* dbg = DBG ? a.dbg : nullptr;
    [[maybe_unused]] const int epi_mode = DBG ? a.epi_mode : 0;

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
        for (int acc = 0; acc < 2; ++acc) {
            ptx::mbar_init(bar_tfull + 8 * acc, 1);            // one tcgen05.commit
            ptx::mbar_init(bar_tempty + 8 * acc, 4 * CG);      // one arrival per epilogue warp (of both CTAs)
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
            // shard rows: NO cache hint (an explicit evict_normal policy is classed
            // "evict_normal_demote" by the L2 and doubled the DRAM reads; evict_first / evict_last
            // were no better -- profiles/README.md)
            const int q_row = q_tile * C::kQTile + static_cast<int>(rank) * kRowsPerCta;
            int stage = 0;
            uint32_t phase = 0;
            [[maybe_unused]] long long t_wait = 0;
            [[maybe_unused]] const long long t_begin = DBG ? clock64() : 0;
            for (int i = 0; i < my_tiles; ++i) {
                const int d_row = (group + i * n_groups) * kTileN + static_cast<int>(rank) * C::kBRows;
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
                    [[maybe_unused]] const long long w0 = DBG ? clock64() : 0;
                    ptx::mbar_wait(bar_empty + 8 * stage, phase ^ 1u);
                    if constexpr (DBG) t_wait += clock64() - w0;
                    const uint32_t sa = base + stage * C::kStageBytes;
                    if constexpr (CG == 1) {
                        const uint32_t fb = bar_full + 8 * stage;
                        ptx::mbar_expect_tx(fb, C::kStageBytes);
                        ptx::tma_load_3d_hint(sa, &tmap_q, kc * kChunkK, q_plane, q_row, fb, pol_q);
                        ptx::tma_load_3d(sa + kABytes, &tmap_d, kc * kChunkK, d_plane, d_row, fb);
                    } else {
                        // both CTAs' bytes are counted on the LEADER's barrier
                        if (rank == 0) ptx::mbar_expect_tx(bar_full + 8 * stage, 2 * C::kStageBytes);
                        const uint32_t fb = ptx::mapa(bar_full + 8 * stage, 0);
                        ptx::tma_load_3d_cg2(sa, &tmap_q, kc * kChunkK, q_plane, q_row, fb, pol_q);
                        ptx::tma_load_3d_cg2_nohint(sa + kABytes, &tmap_d, kc * kChunkK, d_plane, d_row, fb);
                    }
                    if (++stage == kStages) { stage = 0; phase ^= 1u; }
                }
            }
            if constexpr (DBG) {
                if (dbg) {
                    dbg[blockIdx.x * 40 + 0] = clock64() - t_begin;
                    dbg[blockIdx.x * 40 + 1] = t_wait;
                }
            }
        }
    } else if (warp == 1) {
        // -------------------------------------------------------------- MMA issuer
        if (lane == 0 && rank == 0) {
            int stage = 0;
            uint32_t phase = 0;
            [[maybe_unused]] long long t_wfull = 0, t_wtempty = 0;
            [[maybe_unused]] const long long t_begin = DBG ? clock64() : 0;
            for (int i = 0; i < my_tiles; ++i) {
                const int acc = i & 1;
                const uint32_t acc_phase = (i >> 1) & 1;
                [[maybe_unused]] const long long w0 = DBG ? clock64() : 0;
                if constexpr (DBG) {
                    if (dbg && (i == 8 || i == 64 || i == 200)) {            // skew between the units of a group
                        unsigned long long gt;
                        asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(gt));
                        dbg[blockIdx.x * 40 + (i == 8 ? 5 : i == 64 ? 6 : 7)] = gt;
                    }
                }
                ptx::mbar_wait(bar_tempty + 8 * acc, acc_phase ^ 1u);   // epilogue(s) drained it
                if constexpr (DBG) t_wtempty += clock64() - w0;
                ptx::tc_fence_after();
                const uint32_t tmem_d = tmem_base + acc * kTileN;
                for (int kc = 0; kc < kNumChunks * passes; ++kc) {
                    [[maybe_unused]] const long long w1 = DBG ? clock64() : 0;
                    ptx::mbar_wait(bar_full + 8 * stage, phase);        // TMA bytes have landed
                    if constexpr (DBG) t_wfull += clock64() - w1;
                    ptx::tc_fence_after();
                    const uint32_t sa = base + stage * C::kStageBytes;
                    const uint64_t da = make_sw128_desc(sa);
                    const uint64_t db = make_sw128_desc(sa + kABytes);
#pragma unroll
                    for (int k4 = 0; k4 < kChunkK / kUmmaK; ++k4) {
                        // advance 16 elements = 32 bytes inside the swizzle row: +2 in 16-B units
                        ptx::umma_f16<CG>(tmem_d, da + 2 * k4, db + 2 * k4, a.idesc,
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
            if constexpr (DBG) {
                if (dbg) {
                    dbg[blockIdx.x * 40 + 2] = clock64() - t_begin;
                    dbg[blockIdx.x * 40 + 3] = t_wfull;
                    dbg[blockIdx.x * 40 + 4] = t_wtempty;
                }
            }
        }
    } else if (warp == 6) {
        // ---------------------------------------------------------- threshold warp
        if constexpr (LOG)
            threshold_warp_log<R>(C::kLogCap, a.ws_lists, a.ws_counts, a.ws_tau, sm + C::kOffThr, C::kThrSlots, b,
                                  n_qt * C::kQTile, a.gpad, k, n_groups, q_tile * C::kQTile, C::kQTile,
                                  group * CG + static_cast<int>(rank), n_groups * CG, epi_done, lane);
        else if constexpr (!TOP1)                            // k = 1 keeps no lists
            threshold_warp<R>(a.ws_lists, a.ws_tau, b, n_qt * C::kQTile, k, n_groups, q_tile * C::kQTile,
                              C::kQTile, group * CG + static_cast<int>(rank), n_groups * CG, epi_done, lane);
    } else {
        // ---------------------------------------------------------------- epilogue
        const int quarter = warp & 3;                                    // TMEM lanes 32q..32q+31
        const int row0 = q_tile * C::kQTile + static_cast<int>(rank) * kRowsPerCta + quarter * 32;
        const bool row_valid = row0 + lane < b;
        const int b_pad = n_qt * C::kQTile;
        uint32_t* wtau = a.ws_tau + row0;
        [[maybe_unused]] long long t_wtfull = 0, t_ld = 0;
        [[maybe_unused]] const long long t_begin = DBG ? clock64() : 0;
        [[maybe_unused]] unsigned d_slow = 0, d_flush = 0;
        [[maybe_unused]] long long d_tflush = 0;

        if constexpr (TOP1) {
            uint64_t* wlists = a.ws_lists + (static_cast<size_t>(group) * b_pad + row0) * L;
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
        } else if constexpr (LOG) {
            // ------------------------------------------------------ k > 32: candidate logs
            constexpr int CAP = C::kLogCap;
            uint64_t* wlog = a.ws_lists + (static_cast<size_t>(group) * b_pad + row0) * CAP;
            uint32_t* wcount = a.ws_counts + static_cast<size_t>(row0) * a.gpad + group;
            LogState st;
            st.tau_l = __int_as_float(0xff800000);
            st.tau_g = 0u;
            st.thr = st.tau_l;
            st.cnt = 0u;
            st.epoch = 0u;
            st.n_slow = st.n_compact = 0u;
            for (int i = 0; i < my_tiles; ++i) {
                const int t = group + i * n_groups;
                const int acc = i & 1;
                const uint32_t acc_phase = (i >> 1) & 1;
                const uint32_t g = __ldcg(wtau + lane);                    // refresh the shared bound while waiting
                [[maybe_unused]] const long long w0 = DBG ? clock64() : 0;
                ptx::mbar_wait(bar_tfull + 8 * acc, acc_phase);
                if constexpr (DBG) t_wtfull += clock64() - w0;
                ptx::tc_fence_after();
                if (g > st.tau_g) {
                    st.tau_g = g;
                    st.thr = thr_of(st.tau_l, st.tau_g);
                }
                const uint32_t taddr = tmem_base + (static_cast<uint32_t>(quarter * 32) << 16) + acc * kTileN;
                if (i == 0 && a.boot_j > 0 && epi_mode == 0) {
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
                    log_boot_exchange(top, a.boot_j, a.boot_m, row_valid,
                                      a.ws_boot + static_cast<size_t>(row0 + lane) * a.gpad, group, n_groups,
                                      a.ws_arrive + (row0 >> 5), st, lane);
                }
                uint32_t g_next = __ldcg(wtau + lane);
#pragma unroll 1
                for (int c = 0; c < kTileN / 32; ++c) {
                    if (epi_mode == 2) break;                                 // diagnostics: MMA + TMA only
                    if ((c & 3) == 0) {                                       // pick the bound up twice per tile
                        if (g_next > st.tau_g) {
                            st.tau_g = g_next;
                            st.thr = thr_of(st.tau_l, st.tau_g);
                        }
                        g_next = __ldcg(wtau + lane);
                    }
                    uint32_t v[32];
                    [[maybe_unused]] const long long l0 = DBG ? clock64() : 0;
                    ptx::tmem_ld_32x32(taddr + c * 32, v);
                    ptx::tmem_wait_ld();
                    if constexpr (DBG) t_ld += clock64() - l0;
                    if (epi_mode == 1) {                                      // diagnostics: TMEM reads only
                        asm volatile("" ::"r"(v[0]), "r"(v[31]));
                        continue;
                    }
                    process_strip_log<R>(v, static_cast<uint32_t>(t) * kTileN + c * 32, n, row_valid, st, wlog,
                                         wcount, a.gpad, wtau, k, lane);
                }
                ptx::tc_fence_before();
                __syncwarp();
                if (lane == 0) {
                    if constexpr (CG == 1) ptx::mbar_arrive(bar_tempty + 8 * acc);
                    else ptx::mbar_arrive_cluster(bar_tempty + 8 * acc, 0);   // the leader's barrier
                }
            }
            if constexpr (DBG) { d_slow = st.n_slow; d_flush = st.n_compact; }
        } else {
        uint64_t* wlists = a.ws_lists + (static_cast<size_t>(group) * b_pad + row0) * L;
        constexpr int LW = C::kListWidth;
        uint64_t* wbuf = reinterpret_cast<uint64_t*>(sm + C::kOffBuf) + quarter * 32 * kBufStride;
        uint64_t* slists = reinterpret_cast<uint64_t*>(sm + C::kOffLists) + quarter * 32 * LW;
        for (int i = lane; i < 32 * LW; i += 32) slists[i] = 0ull;
        __syncwarp();
        EpiState st;
        st.tau_l = __int_as_float(0xff800000);
        st.tau_g = 0u;
        st.thr = st.tau_l;
        st.cnt = 0;
        st.n_slow = st.n_flush = st.n_cols = 0u;
        st.t_flush = 0;
        for (int i = 0; i < my_tiles; ++i) {
            const int t = group + i * n_groups;
            const int acc = i & 1;
            const uint32_t acc_phase = (i >> 1) & 1;
            // refresh the shared bound while waiting for the accumulator
            const uint32_t g = __ldcg(wtau + lane);
            [[maybe_unused]] const long long w0 = DBG ? clock64() : 0;
            ptx::mbar_wait(bar_tfull + 8 * acc, acc_phase);
            [[maybe_unused]] const long long w1 = DBG ? clock64() : 0;
            if constexpr (DBG) t_wtfull += w1 - w0;
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
                [[maybe_unused]] const long long l0 = DBG ? clock64() : 0;
                ptx::tmem_ld_32x32(taddr + c * 32, v);
                ptx::tmem_wait_ld();
                if constexpr (DBG) t_ld += clock64() - l0;
                if (epi_mode == 1) {                                      // diagnostics: TMEM reads only
                    asm volatile("" ::"r"(v[0]), "r"(v[31]));
                    continue;
                }
                if (!NARROW && i == 0 && k > 16 && c == 0) {              // bootstrap for 16 < k <= 32
                    direct_fill_strip(v, static_cast<uint32_t>(t) * kTileN, n, row_valid, slists, lane);
                    sort_filled_lists(st, slists, wlists, wtau, k, lane);
                    continue;
                }
                process_strip<R, LW, DBG>(v, static_cast<uint32_t>(t) * kTileN + c * 32, n, row_valid, st,
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
            [[maybe_unused]] const long long w2 = DBG ? clock64() : 0;
            const unsigned pending = __ballot_sync(kFull, st.cnt > 0);
            if (pending) flush_lanes<R, LW, DBG>(pending, st, wbuf, slists, wlists, wtau, k, lane);
            if constexpr (DBG) {
                if (dbg && blockIdx.x == 0 && warp == 2 && lane == 0 && i < 64) {
                    unsigned long long* tr = dbg + gridDim.x * 40 + i * 4;
                    tr[0] = w1 - w0;                 // waited for the accumulator
                    tr[1] = w2 - w1;                 // strips (until the accumulator was released)
                    tr[2] = clock64() - w2;          // tile-end list merges
                    tr[3] = __popc(pending);
                }
            }
        }
        if constexpr (DBG) { d_slow = st.n_slow; d_flush = st.n_flush; d_tflush = st.t_flush; }
        }   // k > 1, lists
        __syncwarp();
        if (lane == 0) atomicAdd(epi_done, 1u);
        if constexpr (DBG) {
            if (dbg && lane == 0) {
                unsigned long long* d = dbg + blockIdx.x * 40 + 8 + (warp - 2) * 6;
                d[0] = clock64() - t_begin;
                d[1] = t_wtfull;
                d[2] = d_tflush;
                d[3] = d_slow;
                d[4] = d_flush;
                d[5] = t_ld;
            }
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

// Log mode: fold the candidate logs of every group into the query's top list.  Same geometry as
// above (one warp per query, or eight warps per query when there are many groups); the count
// words of a warp's groups are fetched with one load, the first chunk of the next log while the
// current one is folded.
template <int R>
__global__ void __launch_bounds__(256)
batched_merge_log_kernel(const uint64_t* __restrict__ ws_logs, const uint32_t* __restrict__ ws_counts,
                         uint32_t* __restrict__ ws_tau, int n_groups, int b, int b_pad, int gpad, int k,
                         int warps_per_query, float* __restrict__ out_score, int64_t* __restrict__ out_idx,
                         int64_t idx_offset) {
    constexpr int L = 32 * R;
    constexpr uint32_t CAP = 64 * R;
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
        uint64_t worst = 0ull;
        for (int gb = sub; gb < n_groups; gb += 32 * warps_per_query) {       // 32 of my groups at a time
            const int gm = gb + lane * warps_per_query;
            uint32_t myc = gm < n_groups ? (__ldcg(ws_counts + static_cast<size_t>(query) * gpad + gm) & 0xffffu) : 0u;
            myc = myc > CAP ? CAP : myc;
            const int n_mine = min(32, (n_groups - gb + warps_per_query - 1) / warps_per_query);
            uint32_t cnt = __shfl_sync(kFull, myc, 0);
            const uint64_t* lp = ws_logs + (static_cast<size_t>(gb) * b_pad + query) * CAP;
            uint64_t nxt = static_cast<uint32_t>(lane) < cnt ? __ldcg(lp + lane) : 0ull;
            for (int j = 0; j < n_mine; ++j) {
                uint64_t key = nxt;
                const uint32_t cnt_cur = cnt;
                const uint64_t* lp_cur = lp;
                if (j + 1 < n_mine) {
                    cnt = __shfl_sync(kFull, myc, j + 1);
                    lp = ws_logs + (static_cast<size_t>(gb + (j + 1) * warps_per_query) * b_pad + query) * CAP;
                    nxt = static_cast<uint32_t>(lane) < cnt ? __ldcg(lp + lane) : 0ull;
                }
                for (uint32_t pos = 0; pos < cnt_cur; pos += 32) {
                    if (pos > 0) key = pos + lane < cnt_cur ? __ldcg(lp_cur + pos + lane) : 0ull;
                    unsigned pass = __ballot_sync(kFull, key > worst);
                    while (pass) {
                        const int l = __ffs(pass) - 1;
                        pass &= pass - 1;
                        const uint64_t cand = shfl_u64(key, l);
                        if (cand > worst) {
                            list.insert(cand, lane);
                            worst = list.worst();
                        }
                    }
                }
            }
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
static inline int r_for_k_batched(int k) { return k <= 32 ? 1 : k <= 64 ? 2 : 4; }

static constexpr int64_t kTauBytes = 4096;        // kQueriesPerLaunch * 4
static constexpr int64_t kArriveBytes = 4096;     // log mode: one arrival counter per 32-query slice
static constexpr int64_t kHdrBytes = kTauBytes + kArriveBytes;

// Workspace: [published bounds 4 KB][arrival counters 4 KB][lists: groups x b_pad x 32 R keys]
// or, in log mode (R > 1): [...][logs: groups x b_pad x 64 R keys][counts b_pad x gpad][boot b_pad x gpad]
int64_t batched_workspace_bytes(int64_t /*n*/, int /*b*/, int k, int sm_count) {
    const int R = r_for_k_batched(k);
    const int64_t per_list = (R > 1 ? 64 : 32) * R * 8;
    const int64_t gpad = (sm_count + 31) & ~31;
    const int64_t words = R > 1 ? 2 * k2::kQueriesPerLaunch * gpad * 4 : 0;
    return kHdrBytes + static_cast<int64_t>(sm_count) * k2::kRowsPerCta * per_list + words;
}

template <int R, int CG, bool TOP1, bool DEEP, bool NARROW, bool DBG>
static int launch_batched_k(const CUtensorMap& tq, const CUtensorMap& td, const K2Args& a, cudaStream_t stream) {
    using C = k2::Cfg<CG, R, TOP1, DEEP, NARROW>;
    // per device and cheap: set on every launch (one process may drive several GPUs)
    cudaError_t e = cudaFuncSetAttribute(topk_batched_kernel<R, CG, TOP1, DEEP, NARROW, DBG>,
                                         cudaFuncAttributeMaxDynamicSharedMemorySize, C::kSmemBytes);
    if (e != cudaSuccess) { set_error("topk_batched: smem attribute: %s", cudaGetErrorString(e)); return -2; }
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(static_cast<unsigned>(a.n_groups * a.n_qt * CG));
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
    e = cudaLaunchKernelEx(&cfg, topk_batched_kernel<R, CG, TOP1, DEEP, NARROW, DBG>, tq, td, a);
    if (e != cudaSuccess) { set_error("topk_batched: launch: %s", cudaGetErrorString(e)); return -2; }
    return 0;
}

// DBG_OK: this instantiation also exists with the role timers compiled in (sqe_debug_k2_timers);
// the production instantiation carries no diagnostics at all.
template <int R, int CG, bool TOP1, bool DEEP, bool NARROW = false, bool DBG_OK = false>
static int launch_batched_r(const CUtensorMap& tq, const CUtensorMap& td, K2Args a, float* out_score,
                            int64_t* out_idx, int64_t idx_offset, cudaStream_t stream) {
    using C = k2::Cfg<CG, R, TOP1, DEEP, NARROW>;
    a.dbg = reinterpret_cast<unsigned long long*>(g_k2_debug.load());
    a.epi_mode = g_k2_epilogue_mode.load();
    if (a.n_dtiles > 0) {
        int rc;
        if constexpr (DBG_OK) {
            rc = (a.dbg != nullptr || a.epi_mode != 0)
                     ? launch_batched_k<R, CG, TOP1, DEEP, NARROW, true>(tq, td, a, stream)
                     : launch_batched_k<R, CG, TOP1, DEEP, NARROW, false>(tq, td, a, stream);
        } else {
            rc = launch_batched_k<R, CG, TOP1, DEEP, NARROW, false>(tq, td, a, stream);
        }
        if (rc != 0) return rc;
    }
    const int mg = a.n_dtiles > 0 ? a.n_groups : 0;
    const int b = a.b, k = a.k, b_pad = a.n_qt * C::kQTile;
    if constexpr (TOP1) {
        top1_merge_kernel<<<(b + 3) / 4, 128, 0, stream>>>(a.ws_lists, a.ws_tau, mg, b, b_pad, out_score, out_idx,
                                                           idx_offset);
    } else if constexpr (C::kLog) {
        if (mg > 32)
            batched_merge_log_kernel<R><<<b, 256, 0, stream>>>(a.ws_lists, a.ws_counts, a.ws_tau, mg, b, b_pad, a.gpad,
                                                               k, 8, out_score, out_idx, idx_offset);
        else
            batched_merge_log_kernel<R><<<(b + 3) / 4, 128, 0, stream>>>(a.ws_lists, a.ws_counts, a.ws_tau, mg, b, b_pad,
                                                                         a.gpad, k, 1, out_score, out_idx, idx_offset);
    } else if (mg > 32) {
        batched_merge_kernel<R><<<b, 256, 0, stream>>>(a.ws_lists, a.ws_tau, mg, b, b_pad, k, 8, out_score, out_idx,
                                                       idx_offset);
    } else {
        batched_merge_kernel<R><<<(b + 3) / 4, 128, 0, stream>>>(a.ws_lists, a.ws_tau, mg, b, b_pad, k, 1, out_score,
                                                                 out_idx, idx_offset);
    }
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) { set_error("topk_batched: merge launch: %s", cudaGetErrorString(e)); return -2; }
    return 0;
}

template <int CG>
static int launch_batched_cg(const void* D, int dtype, int64_t n, const void* Q, int b, int k,
                             float* out_score, int64_t* out_idx, int64_t idx_offset, void* ws,
                             int sm_count, cudaStream_t stream) {
    using C = k2::Cfg<CG>;
    const int R = r_for_k_batched(k);
    const bool log_mode = R > 1;
    const int64_t per_list = (log_mode ? 64 : 32) * R * 8;     // bytes per (query, group) list / log
    const int n_dtiles = static_cast<int>((n + k2::kTileN - 1) / k2::kTileN);
    const uint32_t fmt = (dtype == 2) ? 0u : 1u;               // F16 = 0; BF16 = 1 (also split bf16)
    const int64_t q_row_bytes = (dtype == 3) ? 2 * kDim * 2 : kDim * 2;
    K2Args a = {};
    a.n = static_cast<uint32_t>(n);
    a.k = k;
    a.n_dtiles = n_dtiles;
    a.passes = (dtype == 3) ? 4 : 1;                           // split bf16: Ql.Dl + Qh.Dl + Ql.Dh + Qh.Dh
    // instruction descriptor: D fp32 [4,6)=1, A fmt [7,10), B fmt [10,13), K-major both,
    // N>>3 [17,23), M>>4 [24,29)  (M = 128 per CTA, 256 for the pair)
    a.idesc = (1u << 4) | (fmt << 7) | (fmt << 10) | (static_cast<uint32_t>(k2::kTileN >> 3) << 17) |
              (static_cast<uint32_t>(C::kQTile >> 4) << 24);
    a.ws_tau = static_cast<uint32_t*>(ws);
    a.ws_arrive = reinterpret_cast<uint32_t*>(static_cast<char*>(ws) + kTauBytes);
    a.ws_lists = reinterpret_cast<uint64_t*>(static_cast<char*>(ws) + kHdrBytes);
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
        if (n_groups > k2::kMaxGroups) n_groups = k2::kMaxGroups;
        const int64_t b_pad = static_cast<int64_t>(n_qt) * C::kQTile;
        const int64_t lists_bytes = static_cast<int64_t>(n_groups) * b_pad * per_list;
        a.b = bc;
        a.n_qt = n_qt;
        a.n_groups = n_groups;
        a.gpad = (n_groups + 31) & ~31;
        int64_t used = kHdrBytes + lists_bytes;
        a.boot_j = a.boot_m = 0;
        if (log_mode) {
            a.ws_counts = reinterpret_cast<uint32_t*>(static_cast<char*>(ws) + used);
            a.ws_boot = a.ws_counts + b_pad * a.gpad;
            used += 2 * b_pad * a.gpad * 4;
            // first-tile bootstrap: the m-th largest over the groups of the j-th best of a group's
            // first d-tile bounds the k-th best when j m >= k; j, m <= 16 (register networks)
            int j = (k + 15) / 16;
            while (j <= 16 && (k + j - 1) / j > n_groups) ++j;
            if (j <= 16 && n_groups >= 2) {
                a.boot_j = j;
                a.boot_m = (k + j - 1) / j;
            }
        }
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
#define SQE_K2_LAUNCH(R_, TOP1_, DBG_DEEP_, DBG_FLAT_)                                                              \
    (deep ? launch_batched_r<R_, CG, TOP1_, (CG == 2), false, DBG_DEEP_>(tq, td, a, os, oi, idx_offset, stream)     \
          : launch_batched_r<R_, CG, TOP1_, false, false, DBG_FLAT_>(tq, td, a, os, oi, idx_offset, stream))
        if (k == 1) rc = SQE_K2_LAUNCH(1, true, false, false);
        else if (CG == 1 && k <= 16)         // narrow smem lists + a fourth stage (see Cfg)
            rc = launch_batched_r<1, CG, false, false, (CG == 1)>(tq, td, a, os, oi, idx_offset, stream);
        else if (R == 1) rc = SQE_K2_LAUNCH(1, false, false, (CG == 2));       // role timers: the b = 1024 headline form
        else if (R == 2) rc = SQE_K2_LAUNCH(2, false, false, false);
        else rc = SQE_K2_LAUNCH(4, false, (CG == 2), false);                   // role timers: the configs[3] form
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
    int cg = g_k2_cta_group.load();
    if (cg == 0) cg = (b > k2::kRowsPerCta) ? 2 : 1;
    if (cg == 2)
        return launch_batched_cg<2>(D, dtype, n, Q, b, k, out_score, out_idx, idx_offset, ws, sm_count, stream);
    return launch_batched_cg<1>(D, dtype, n, Q, b, k, out_score, out_idx, idx_offset, ws, sm_count, stream);
}

}  // namespace sqe
