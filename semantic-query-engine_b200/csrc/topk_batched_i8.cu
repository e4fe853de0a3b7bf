// K2p  batched exact cosine top-k at HALF the bytes and TWICE the tensor rate: an int8 prefilter on
// the 5th-gen tensor cores (tcgen05.mma kind::i8, s32 accumulators in TMEM, TMA-fed) with a
// rigorous error bound, followed by exact rescoring of the few (query, row) pairs the bound cannot
// rule out.  The batched counterpart of K3p (topk_prefilter.cu); same contract as K2/K3 -- it
// replaces the k-NN request of OpenSearchIndexer.search (app/main.py:356-367) for a batch of
// queries and, with k = 1, the scan of lfu_cache_get (app/main.py:73-90) -- and the SAME RESULTS
// AS K3 (topk_gemv.cu) BIT FOR BIT: the exact pass scores a row with K3's fp32 operation order.
//
//   prepare_queries_kernel   raw fp32 queries -> stored unit queries (K1 arithmetic, rounded to the
//                            shard's storage class) + their int8 quantisation q = sq q8 + eq and
//                            the query constants {sq, qe', qn'}
//   topk_batched_i8_kernel   S8 = Q8 . D8^T on the tensor cores (exact integer arithmetic), tile by
//                            tile like K2 (TMA ring -> tcgen05.mma -> two TMEM accumulators -> four
//                            epilogue warps).  With the row constants {sd, eps, nd} of K1q
//                            (sqe_quantize_rows) the epilogue turns an accumulator into
//                                s8 = sd sq acc,   m = qe' nd + qn' eps,   L = s8 - m <= score <= s8 + m = U
//                            (Cauchy-Schwarz, topk_prefilter.cu) where `score` is the fp32 number K3
//                            computes.  tau = the k-th best L is a lower bound of the k-th best score,
//                            so only pairs with U >= tau can be in the result.  tau is not known
//                            until the end; the kernel uses the running, GPU-wide k-th best L (the
//                            candidate-log machinery of K2's k > 32 mode: every pair with U above the
//                            running bound is appended to the (query, group) CANDIDATE log -- write-only
//                            during the scan --, the few pairs whose L itself beats the bound also to
//                            the small BOUND log the threshold warps merge incrementally to publish
//                            the k-th best L) -- a bound that only grows, so the logged set is a superset.
//   batched_rescore_kernel   one CTA per query: every logged row is scored exactly (K3's loads, FMA
//                            chains and butterfly -> K3's bits) and goes through the warp-list /
//                            CTA-merge selection.  A query whose log overflowed (adversarial data:
//                            sorted scores, a non-finite query) is scanned exactly over ALL rows.
//
// Nothing is approximate.  Roofline: HBM for small batches (n * 1040 B per batch instead of
// n * 2048), the int8 tensor pipe (2x the bf16 rate) for large ones.
#include "sqe_k2.cuh"
#include "sqe_rowload.cuh"

namespace sqe {

namespace k2i {
using namespace k2;
constexpr int kChunkI8 = 128;                      // int8 elements per K chunk = one 128-byte swizzle row
constexpr int kNumChunksI8 = kDim / kChunkI8;      // 8
constexpr int kUmmaKI8 = 32;                       // elements per tcgen05.mma kind::i8
constexpr int kLogCapMax = 1024;                   // entries per (query, group) candidate log
constexpr int kSpillCap = 8192;                    // entries per query of the shared spill area (full logs)
#ifndef SQE_I8_META_BUFS
#define SQE_I8_META_BUFS 4
#endif
#ifndef SQE_I8_ASTAT_STAGES
#define SQE_I8_ASTAT_STAGES 4
#endif
#ifndef SQE_I8_MIN_THR_SLOTS
#define SQE_I8_MIN_THR_SLOTS 4
#endif
constexpr int kMetaBufs = SQE_I8_META_BUFS;        // row-constant tiles in flight (see the producer)
constexpr int kMetaBytes = kTileN * 16;            // {sd, eps, nd, 0} per shard row of a d-tile
constexpr float kSlack = 4e-6f;                    // as in topk_prefilter.cu (pf::kSlack, pf::kInflate)
constexpr float kInflate = 1.001f;

template <int CG, int R, bool DEEP>
struct CfgI8 {
    static constexpr int kQTile = kRowsPerCta * CG;
    static constexpr int kBRows = kTileN / CG;
    static constexpr int kBBytes = kBRows * kChunkI8;
    // ASTAT (pairs with several q-tiles in flight, i.e. b > 256): the unit's int8 QUERY tile stays in
    // shared memory for the whole kernel (128 rows x 1024 B = 128 KB per CTA -- half of what a 16-bit
    // tile needs, which is why K2 cannot do this) and only the shard rows stream through the ring.
    // Without it every d-tile re-fetches the query tile from L2: 32 KB per K chunk and CTA = 128 B per
    // SM clock at the int8 rate, three times what the L2 delivers (~43 B per SM clock chip-wide): the
    // main loop alone took 6.8 ms for a floor of 3.6.  With it the ring carries 16 KB per chunk.
    static constexpr bool kAStat = (CG == 2) && !DEEP;
    static constexpr int kAResident = kAStat ? kNumChunksI8 * kABytes : 0;      // 128 KB
    static constexpr int kStageBytes = (kAStat ? 0 : kABytes) + kBBytes;   // 48 KB / 32 KB as for 16-bit operands; 16 KB
    static constexpr int kStages = (CG == 1) ? 4 : (DEEP ? 6 : (R == 1 ? SQE_I8_ASTAT_STAGES : 4));
    static constexpr int kOffStages = kAResident;
    static constexpr int kOffMeta = kOffStages + kStages * kStageBytes;
    static constexpr int kOffXbuf = kOffMeta + kMetaBufs * kMetaBytes;     // per epilogue warp: 32 words of transpose
    static constexpr int kXbufWords = 32 + kTileN;                         // buffer + the tile's 256 row scales, packed
    static constexpr int kOffThr = kOffXbuf + 4 * kXbufWords * 4;
    static constexpr int kThrSlotBytes = 32 * R * 8 + kMaxGroups * 4;
    static constexpr int kBarBytes = 32 * 8 + 16;
    static constexpr int kFree = 227 * 1024 - 1024 - kBarBytes - kOffThr;
    static constexpr int kThrSlots = (kFree / kThrSlotBytes) < 48 ? (kFree / kThrSlotBytes) : 48;
    static constexpr int kOffBar = kOffThr + kThrSlots * kThrSlotBytes;
    static constexpr int kOffTmemPtr = kOffBar + 32 * 8;
    static constexpr int kSmemBytes = kOffTmemPtr + 16 + 1024;
    static_assert(kThrSlots >= (R == 1 ? SQE_I8_MIN_THR_SLOTS : 4) && kSmemBytes <= 227 * 1024, "shared memory budget");
};
}  // namespace k2i

struct K2I8Args {
    uint32_t n;
    int b, k;
    int n_qt, n_groups, n_dtiles;
    uint32_t idesc;
    uint64_t* ws_logs;        // bound logs [group][b_pad][64 R] keys (L, row) with L above the bound when logged
    uint64_t* ws_clogs;       // candidate logs [group][b_pad][log_cap] keys (L, row) with U above the bound
    uint32_t* ws_ccounts;     // [b_pad][gpad] entries per candidate log (written when the scan ends)
    uint32_t* ws_tau;         // published bounds (k-th best L), first 4 KB of the workspace
    uint32_t* ws_arrive;      // bootstrap arrivals per 32-query slice
    uint32_t* ws_counts;      // [b_pad][gpad]
    uint32_t* ws_boot;        // [b_pad][gpad]
    uint32_t* ws_over;        // [b_pad] != 0: log AND spill area of this query overflowed -> exact scan of every row
    uint32_t* ws_spill_cnt;   // [b_pad] entries in the query's spill area
    uint64_t* ws_spill;       // [b_pad][kSpillCap] keys that did not fit their (query, group) log (never zeroed:
                              // read by the exact pass only, after the kernel, up to the count)
    int log_cap;              // entries per candidate log
    int blog_cap;             // entries per bound log (64 R)
    int epi_mode;             // diagnostics (SQE_TUNE_K2_EPILOGUE_MODE): 2 = no epilogue work (results invalid)
    uint32_t* ws_prog;        // [group][n_qt] d-tile progress of the units that share d-tiles (in the arrivals page)
    int window;               // how many d-tiles a unit may run ahead of the slowest sibling (0 = unbounded)
    int gpad;
    int boot_j, boot_m;
    const float4* meta;       // row constants {sd, eps, nd, 0} (sqe_quantize_rows)
    const float4* qmeta;      // query constants {sq, qe', qn', 0} (prepare_queries_kernel)
};

__device__ __forceinline__ void umma_i8(uint32_t tmem_d, uint64_t desc_a, uint64_t desc_b, uint32_t idesc,
                                        uint32_t accumulate, bool pair) {
    if (!pair)
        asm volatile(
            "{\n.reg .pred p;\nsetp.ne.b32 p, %4, 0;\n"
            "tcgen05.mma.cta_group::1.kind::i8 [%0], %1, %2, %3, p;\n}\n" ::"r"(tmem_d),
            "l"(desc_a), "l"(desc_b), "r"(idesc), "r"(accumulate)
            : "memory");
    else
        asm volatile(
            "{\n.reg .pred p;\nsetp.ne.b32 p, %4, 0;\n"
            "tcgen05.mma.cta_group::2.kind::i8 [%0], %1, %2, %3, p;\n}\n" ::"r"(tmem_d),
            "l"(desc_a), "l"(desc_b), "r"(idesc), "r"(accumulate)
            : "memory");
}

// plain (non-tensor) bulk copy global -> this CTA's shared memory, bytes signalled on a local mbarrier
__device__ __forceinline__ void bulk_copy_g2s(uint32_t dst, const void* src, uint32_t bytes, uint32_t bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(dst),
                 "l"(src), "r"(bytes), "r"(bar)
                 : "memory");
}

struct I8State {
    float thr;        // a pair is logged iff !(U <= thr)
    uint32_t tau_g;   // best published bound on the k-th best L (orderable u32), 0 = none
    uint32_t cnt;     // entries in this thread's candidate log
    uint32_t bcnt;    // entries in this thread's bound log
    bool dead;        // log and spill area overflowed: the query is flagged for the exact scan, nothing more is logged
    float sq, qe, qn; // query constants
};

// Per-warp pointers of the epilogue's logs: lane w's log / counter is base + w * stride.
struct I8Logs {
    uint64_t* clog;       // candidate logs of this warp's 32 queries (stride log_cap)
    uint64_t* blog;       // bound logs (stride blog_cap)
    uint32_t* bcount;     // bound-log count words (stride gpad)
    uint64_t* spill;      // spill areas (stride kSpillCap)
    uint32_t* spill_cnt;  // (stride 1)
    uint32_t* over;       // (stride 1)
    uint32_t log_cap, blog_cap, gpad;
    uint32_t* xbuf;       // 32 words of shared memory per warp: one lane's accumulators, transposed
    const float* sds;     // the d-tile's 256 row scales sd, packed (per warp; written in the tile prologue)
};

// One 32-column strip: v[j] = the s32 accumulator of (this thread's query, row col0 + j).
// mt = the row constants of this d-tile in shared memory; m_max = the margin with the tile's
// largest nd and eps (>= every column's margin: the expression is monotone in both).
//
// Fast path (per thread = per query): can any column's upper bound U reach thr?  With a margin of
// ~half a score sigma that is true for some lane of the warp in about every second strip, so the
// slow path is TRANSPOSED: a lane that wants hands its 32 accumulators over through shared
// memory and the 32 lanes evaluate one column each (exact bounds, one conflict-free LDS.128 of
// the row constants), then append the passing columns to that query's logs in parallel.
__device__ __forceinline__ void process_strip_i8(const uint32_t (&v)[32], uint32_t col0, int c_in_tile, uint32_t n,
                                                 bool row_valid, I8State& st, const float4* mt, float m_max,
                                                 const I8Logs& lg, int lane) {
    const float4* mc = mt + c_in_tile * 32;
    float r[32];
#pragma unroll
    for (int j = 0; j < 32; j += 4) {                            // acc * sd, four row scales per (broadcast) LDS.128
        const float4 a4 = *reinterpret_cast<const float4*>(lg.sds + c_in_tile * 32 + j);
        r[j] = static_cast<float>(static_cast<int>(v[j])) * a4.x;
        r[j + 1] = static_cast<float>(static_cast<int>(v[j + 1])) * a4.y;
        r[j + 2] = static_cast<float>(static_cast<int>(v[j + 2])) * a4.z;
        r[j + 3] = static_cast<float>(static_cast<int>(v[j + 3])) * a4.w;
    }
    if (col0 + 32u > n) {                                       // ragged last d-tile (warp-uniform)
#pragma unroll
        for (int j = 0; j < 32; ++j)
            if (col0 + j >= n) r[j] = __int_as_float(0xff800000);
    }
    float t[11];
#pragma unroll
    for (int i = 0; i < 10; ++i) t[i] = fmax3(r[3 * i], r[3 * i + 1], r[3 * i + 2]);
    t[10] = fmaxf(r[30], r[31]);
    const float u0 = fmax3(t[0], t[1], t[2]), u1 = fmax3(t[3], t[4], t[5]);
    const float u2 = fmax3(t[6], t[7], t[8]), u3 = fmaxf(t[9], t[10]);
    const float rmax = fmaxf(fmaxf(u0, u1), fmaxf(u2, u3));
    // no column of the strip can reach thr if even the largest s8 plus the largest margin stays below
    // it; (sq rmax) differs from the slow path's (sd sq) acc by roundings only: the relative and
    // absolute slack covers them.  Negated comparison: a NaN (non-finite query or row constants)
    // goes to the slow path.
    float hi = st.sq * rmax;
    hi = hi + fabsf(hi) * 2e-6f + 2e-7f;
    const bool want = row_valid && !st.dead && !(hi <= st.thr - m_max);
    unsigned todo = __ballot_sync(kFull, want);
    while (todo) {
        const int w = __ffs(todo) - 1;
        todo &= todo - 1;
        if (lane == w) {
#pragma unroll
            for (int j = 0; j < 32; j += 4)
                *reinterpret_cast<uint4*>(lg.xbuf + j) = make_uint4(v[j], v[j + 1], v[j + 2], v[j + 3]);
        }
        __syncwarp();
        const int acc = static_cast<int>(lg.xbuf[lane]);        // column `lane` of query w
        const float sq = __shfl_sync(kFull, st.sq, w);
        const float qe = __shfl_sync(kFull, st.qe, w);
        const float qn = __shfl_sync(kFull, st.qn, w);
        const float thr = __shfl_sync(kFull, st.thr, w);
        const uint32_t cnt = __shfl_sync(kFull, st.cnt, w);
        const uint32_t bcnt = __shfl_sync(kFull, st.bcnt, w);
        const float4 m4 = mc[lane];
        const float s8 = (m4.x * sq) * static_cast<float>(acc);
        const float m = fmaf(qe, m4.z, qn * m4.y) + 1e-30f;
        const float U = s8 + m;
        const float Lb = s8 - m;
        const uint32_t col = col0 + lane;
        const bool pass = col < n && !(U <= thr);
        const unsigned pm = __ballot_sync(kFull, pass);
        const unsigned bm = __ballot_sync(kFull, pass && Lb > thr);
        const unsigned below = (1u << lane) - 1u;
        bool over = false;
        if (pass) {
            const uint64_t key = make_key(Lb, col);              // keyed by L (NaN -> -inf)
            const uint32_t slot = cnt + __popc(pm & below);
            if (slot < lg.log_cap) {
                __stcg(lg.clog + static_cast<size_t>(w) * lg.log_cap + slot, key);
            } else {
                // this candidate log is full (a loose bound: clustered or sorted data): the pair
                // goes to the query's spill area, shared by all groups
                const uint32_t sp = atomicAdd(lg.spill_cnt + w, 1u);
                if (sp < static_cast<uint32_t>(k2i::kSpillCap))
                    __stcg(lg.spill + static_cast<size_t>(w) * k2i::kSpillCap + sp, key);
                else over = true;
            }
            // a lower bound that itself beats the running bound can raise it: bound log (a full one
            // just stops feeding the threshold warps; the bound stays valid)
            if ((bm >> lane) & 1u) {
                const uint32_t bslot = bcnt + __popc(bm & below);
                if (bslot < lg.blog_cap) __stcg(lg.blog + static_cast<size_t>(w) * lg.blog_cap + bslot, key);
            }
        }
        over = __any_sync(kFull, over);
        if (lane == w) {
            st.cnt = min(cnt + __popc(pm), lg.log_cap);
            const uint32_t nb = min(bcnt + __popc(bm), lg.blog_cap);
            if (nb != st.bcnt) {
                st.bcnt = nb;
                st_relaxed_gpu(lg.bcount + static_cast<size_t>(w) * lg.gpad, nb);   // epoch 0: never compacted
            }
            if (over) {
                st.dead = true;
                st_relaxed_gpu(lg.over + w, 1u);
            }
        }
        __syncwarp();                                           // xbuf is reused by the next lane
    }
}

// First d-tile: the 16 largest of the 64 maxima of 4 consecutive columns' LOWER bounds L.
__device__ __forceinline__ void bootstrap_strip_i8(const uint32_t (&v)[32], uint32_t col0, int c_in_tile, uint32_t n,
                                                   const I8State& st, const float4* mt, float (&top)[16]) {
    const float4* mc = mt + c_in_tile * 32;
#pragma unroll
    for (int g = 0; g < 8; ++g) {
        float x = __int_as_float(0xff800000);
#pragma unroll
        for (int e = 0; e < 4; ++e) {
            const int j = 4 * g + e;
            const float4 m4 = mc[j];
            const float s8 = (m4.x * st.sq) * static_cast<float>(static_cast<int>(v[j]));
            const float m = fmaf(st.qe, m4.z, st.qn * m4.y) + 1e-30f;
            const float L = s8 - m;
            x = fmaxf(x, (col0 + j < n && L == L) ? L : __int_as_float(0xff800000));
        }
#pragma unroll
        for (int i = 0; i < 16; ++i) {
            const float hi = fmaxf(top[i], x);
            x = fminf(top[i], x);
            top[i] = hi;
        }
    }
}

template <int R, int CG, bool DEEP>
__global__ void __launch_bounds__(k2::kThreads, 1)
topk_batched_i8_kernel(const __grid_constant__ CUtensorMap tmap_q, const __grid_constant__ CUtensorMap tmap_d,
                       const K2I8Args a) {
    using namespace k2i;
    using C = CfgI8<CG, R, DEEP>;
    constexpr int kStages = C::kStages;
    extern __shared__ uint8_t smem_raw[];
    const uint32_t raw_addr = ptx::smem_u32(smem_raw);
    const uint32_t base = (raw_addr + 1023u) & ~1023u;
    uint8_t* sm = smem_raw + (base - raw_addr);

    const uint32_t n = a.n;
    const int b = a.b, k = a.k, n_qt = a.n_qt, n_groups = a.n_groups, n_dtiles = a.n_dtiles;
    const int warp = __shfl_sync(kFull, static_cast<int>(threadIdx.x >> 5), 0);
    const int lane = threadIdx.x & 31;
    const uint32_t rank = (CG == 2) ? ptx::cluster_ctarank() : 0u;
    const int unit = blockIdx.x / CG;
    const int q_tile = unit % n_qt;
    const int group = unit / n_qt;
    const int my_tiles = (group < n_dtiles) ? (n_dtiles - group + n_groups - 1) / n_groups : 0;

    const uint32_t bar_full = base + C::kOffBar;               // [kStages]
    const uint32_t bar_empty = bar_full + 8 * kStages;         // [kStages]
    const uint32_t bar_tfull = bar_empty + 8 * kStages;        // [2]
    const uint32_t bar_tempty = bar_tfull + 16;                // [2]
    const uint32_t bar_meta = bar_tempty + 16;                 // [kMetaBufs]
    const uint32_t bar_a = bar_meta + 8 * kMetaBufs;           // ASTAT: the resident query tile has landed
    constexpr bool ASTAT = C::kAStat;
    const uint32_t stage0 = base + C::kOffStages;
    uint32_t* tmem_ptr_smem = reinterpret_cast<uint32_t*>(sm + C::kOffTmemPtr);
    uint32_t* epi_done = tmem_ptr_smem + 1;

    if (warp == 0 && lane == 0) {
        ptx::prefetch_tensormap(&tmap_q);
        ptx::prefetch_tensormap(&tmap_d);
        for (int s = 0; s < kStages; ++s) {
            ptx::mbar_init(bar_full + 8 * s, 1);
            ptx::mbar_init(bar_empty + 8 * s, 1);
        }
        for (int acc = 0; acc < 2; ++acc) {
            ptx::mbar_init(bar_tfull + 8 * acc, 1);
            ptx::mbar_init(bar_tempty + 8 * acc, 4 * CG);
        }
        for (int m = 0; m < kMetaBufs; ++m) ptx::mbar_init(bar_meta + 8 * m, 1);
        ptx::mbar_init(bar_a, 1);
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
            const uint64_t pol_q = ptx::policy_evict_last();   // the query tile is re-read for every d-tile: keep it in L2
            const int q_row = q_tile * C::kQTile + static_cast<int>(rank) * kRowsPerCta;
            int stage = 0;
            uint32_t phase = 0;
            // The units of a group (one per q-tile) read the SAME d-tiles and should read them from L2
            // once one of them has fetched them from HBM.  This kernel is bound by its epilogue, whose
            // pace differs from unit to unit (slow-path frequency), so without a bound the units drift
            // apart by hundreds of tiles and every d-tile is fetched from HBM up to n_qt times (ncu:
            // 2.3 x the int8 shard).  A unit therefore publishes the tile it is about to load and does
            // not run more than `window` tiles ahead of its slowest sibling; the group finishes when
            // its slowest unit does, so waiting costs nothing.
            uint32_t* my_prog = a.ws_prog + group * n_qt;
            const bool gate = (rank == 0) && (n_qt > 1) && (a.window > 0);
            if constexpr (ASTAT) {
                if (my_tiles > 0) {                              // the query tile, once: eight 16 KB chunks per CTA
                    if (rank == 0) ptx::mbar_expect_tx(bar_a, 2 * C::kAResident);
                    const uint32_t ab = ptx::mapa(bar_a, 0);
                    for (int kc = 0; kc < kNumChunksI8; ++kc)
                        ptx::tma_load_3d_cg2(base + kc * kABytes, &tmap_q, kc * kChunkI8, 0, q_row, ab, pol_q);
                }
            }
            for (int i = 0; i < my_tiles; ++i) {
                const int t = group + i * n_groups;
                const int d_row = t * kTileN + static_cast<int>(rank) * C::kBRows;
                if (gate) {
                    st_relaxed_gpu(my_prog + q_tile, static_cast<uint32_t>(i + 1));
                    const long long t0 = clock64();
                    while (true) {
                        uint32_t lo = 0xffffffffu;
                        for (int u = 0; u < n_qt; ++u) lo = min(lo, ld_relaxed_gpu(my_prog + u));
                        if (lo + static_cast<uint32_t>(a.window) >= static_cast<uint32_t>(i + 1)) break;
                        if (clock64() - t0 > 2000000000LL) break;      // never hang on a sibling: just stop waiting
                        __nanosleep(200);
                    }
                }
                for (int kc = 0; kc < kNumChunksI8; ++kc) {
                    ptx::mbar_wait(bar_empty + 8 * stage, phase ^ 1u);
                    if (kc == 0) {
                        // the row constants of this d-tile, for THIS CTA's epilogue (both CTAs of a pair
                        // need all 256 rows).  Buffer i % 4: the producer is never more than three tiles
                        // ahead of the epilogue (ring depth < 1 tile, two accumulators), so the buffer's
                        // previous user (tile i - 4) has been consumed.
                        const uint32_t rows = min(static_cast<uint32_t>(kTileN), n - static_cast<uint32_t>(t) * kTileN);
                        const uint32_t mb = static_cast<uint32_t>(i) % kMetaBufs;
                        ptx::mbar_expect_tx(bar_meta + 8 * mb, rows * 16u);
                        bulk_copy_g2s(base + C::kOffMeta + mb * kMetaBytes, a.meta + static_cast<size_t>(t) * kTileN,
                                      rows * 16u, bar_meta + 8 * mb);
                    }
                    const uint32_t sa = stage0 + stage * C::kStageBytes;
                    if constexpr (CG == 1) {
                        const uint32_t fb = bar_full + 8 * stage;
                        ptx::mbar_expect_tx(fb, C::kStageBytes);
                        ptx::tma_load_3d_hint(sa, &tmap_q, kc * kChunkI8, 0, q_row, fb, pol_q);
                        ptx::tma_load_3d(sa + kABytes, &tmap_d, kc * kChunkI8, 0, d_row, fb);
                    } else if constexpr (ASTAT) {
                        if (rank == 0) ptx::mbar_expect_tx(bar_full + 8 * stage, 2 * C::kStageBytes);
                        const uint32_t fb = ptx::mapa(bar_full + 8 * stage, 0);
                        ptx::tma_load_3d_cg2_nohint(sa, &tmap_d, kc * kChunkI8, 0, d_row, fb);      // shard rows only
                    } else {
                        if (rank == 0) ptx::mbar_expect_tx(bar_full + 8 * stage, 2 * C::kStageBytes);
                        const uint32_t fb = ptx::mapa(bar_full + 8 * stage, 0);
                        ptx::tma_load_3d_cg2(sa, &tmap_q, kc * kChunkI8, 0, q_row, fb, pol_q);
                        ptx::tma_load_3d_cg2_nohint(sa + kABytes, &tmap_d, kc * kChunkI8, 0, d_row, fb);
                    }
                    if (++stage == kStages) { stage = 0; phase ^= 1u; }
                }
            }
            if (gate) st_relaxed_gpu(my_prog + q_tile, 0xffffffffu);    // done: never hold the others back
        }
    } else if (warp == 1) {
        // -------------------------------------------------------------- MMA issuer
        if (lane == 0 && rank == 0) {
            int stage = 0;
            uint32_t phase = 0;
            if constexpr (ASTAT) {
                if (my_tiles > 0) {
                    ptx::mbar_wait(bar_a, 0);                   // the resident query tile (both CTAs' halves)
                    ptx::tc_fence_after();
                }
            }
            for (int i = 0; i < my_tiles; ++i) {
                const int acc = i & 1;
                const uint32_t acc_phase = (i >> 1) & 1;
                ptx::mbar_wait(bar_tempty + 8 * acc, acc_phase ^ 1u);
                ptx::tc_fence_after();
                const uint32_t tmem_d = tmem_base + acc * kTileN;
                for (int kc = 0; kc < kNumChunksI8; ++kc) {
                    ptx::mbar_wait(bar_full + 8 * stage, phase);
                    ptx::tc_fence_after();
                    const uint32_t sa = stage0 + stage * C::kStageBytes;
                    const uint64_t da = make_sw128_desc(ASTAT ? base + kc * kABytes : sa);
                    const uint64_t db = make_sw128_desc(ASTAT ? sa : sa + kABytes);
#pragma unroll
                    for (int k4 = 0; k4 < kChunkI8 / kUmmaKI8; ++k4)     // 32 int8 = 32 bytes: +2 in 16-B units
                        umma_i8(tmem_d, da + 2 * k4, db + 2 * k4, a.idesc, (kc | k4) != 0 ? 1u : 0u, CG == 2);
                    if constexpr (CG == 1) ptx::umma_commit(bar_empty + 8 * stage);
                    else ptx::umma_commit_cg2(bar_empty + 8 * stage, 0x3);
                    if (++stage == kStages) { stage = 0; phase ^= 1u; }
                }
                if constexpr (CG == 1) ptx::umma_commit(bar_tfull + 8 * acc);
                else ptx::umma_commit_cg2(bar_tfull + 8 * acc, 0x3);
            }
        }
    } else if (warp == 6) {
        // ---------------------------------------------------------- threshold warp
        threshold_warp_log<R>(a.blog_cap, a.ws_logs, a.ws_counts, a.ws_tau, sm + C::kOffThr, C::kThrSlots, b,
                              n_qt * C::kQTile, a.gpad, k, n_groups, q_tile * C::kQTile, C::kQTile,
                              group * CG + static_cast<int>(rank), n_groups * CG, epi_done, lane, 1);
    } else {
        // ---------------------------------------------------------------- epilogue
        const int quarter = warp & 3;
        const int row0 = q_tile * C::kQTile + static_cast<int>(rank) * kRowsPerCta + quarter * 32;
        const int row = row0 + lane;
        const bool row_valid = row < b;
        const int b_pad = n_qt * C::kQTile;
        uint32_t* wtau = a.ws_tau + row0;
        I8Logs lg;
        lg.log_cap = static_cast<uint32_t>(a.log_cap);
        lg.blog_cap = static_cast<uint32_t>(a.blog_cap);
        lg.gpad = static_cast<uint32_t>(a.gpad);
        lg.clog = a.ws_clogs + (static_cast<size_t>(group) * b_pad + row0) * lg.log_cap;
        lg.blog = a.ws_logs + (static_cast<size_t>(group) * b_pad + row0) * lg.blog_cap;
        lg.bcount = a.ws_counts + static_cast<size_t>(row0) * a.gpad + group;
        lg.over = a.ws_over + row0;
        lg.spill = a.ws_spill + static_cast<size_t>(row0) * kSpillCap;
        lg.spill_cnt = a.ws_spill_cnt + row0;
        lg.xbuf = reinterpret_cast<uint32_t*>(sm + C::kOffXbuf) + quarter * C::kXbufWords;
        float* sds = reinterpret_cast<float*>(lg.xbuf + 32);
        lg.sds = sds;
        I8State st;
        st.tau_g = 0u;
        st.thr = __int_as_float(0xff800000);
        st.cnt = 0u;
        st.bcnt = 0u;
        st.dead = false;
        const float4 qm = row_valid ? a.qmeta[row] : make_float4(0.f, 0.f, 0.f, 0.f);
        st.sq = qm.x;
        st.qe = qm.y;
        st.qn = qm.z;
        const float ninf = __int_as_float(0xff800000);
        for (int i = 0; i < my_tiles; ++i) {
            const int t = group + i * n_groups;
            const int acc = i & 1;
            const uint32_t acc_phase = (i >> 1) & 1;
            const uint32_t mb = static_cast<uint32_t>(i) % kMetaBufs;
            const uint32_t g = __ldcg(wtau + lane);
            ptx::mbar_wait(bar_meta + 8 * mb, (static_cast<uint32_t>(i) / kMetaBufs) & 1u);
            const float4* mt = reinterpret_cast<const float4*>(sm + C::kOffMeta + mb * kMetaBytes);
            // largest nd and eps of the tile's valid rows -> the largest margin of this thread's query
            float bmax = 0.f, cmax = 0.f;
            bool odd = false;                                       // a NaN constant: every strip takes the slow path
#pragma unroll
            for (int j = 0; j < kTileN / 32; ++j) {
                const uint32_t col = static_cast<uint32_t>(t) * kTileN + j * 32 + lane;
                const float4 m4r = mt[j * 32 + lane];
                sds[j * 32 + lane] = col < n ? m4r.x : 0.f;
                if (col < n) {
                    const float4 m4 = m4r;
                    bmax = fmaxf(bmax, m4.z);
                    cmax = fmaxf(cmax, m4.y);
                    odd |= !(m4.z == m4.z) || !(m4.y == m4.y) || !(m4.x == m4.x);
                }
            }
#pragma unroll
            for (int d = 16; d >= 1; d >>= 1) {
                bmax = fmaxf(bmax, __shfl_xor_sync(kFull, bmax, d));
                cmax = fmaxf(cmax, __shfl_xor_sync(kFull, cmax, d));
            }
            odd = __any_sync(kFull, odd);
            __syncwarp();                                        // the packed row scales are visible to the whole warp
            const float m_max = odd ? __int_as_float(0x7fc00000) : fmaf(st.qe, bmax, st.qn * cmax) + 1e-30f;
            ptx::mbar_wait(bar_tfull + 8 * acc, acc_phase);
            ptx::tc_fence_after();
            if (g > st.tau_g) {
                st.tau_g = g;
                st.thr = thr_of(ninf, st.tau_g);
            }
            const uint32_t taddr = tmem_base + (static_cast<uint32_t>(quarter * 32) << 16) + acc * kTileN;
            if (i == 0 && a.boot_j > 0) {
                float top[16];
#pragma unroll
                for (int j = 0; j < 16; ++j) top[j] = ninf;
#pragma unroll 1
                for (int c = 0; c < kTileN / 32; ++c) {
                    uint32_t v[32];
                    ptx::tmem_ld_32x32(taddr + c * 32, v);
                    ptx::tmem_wait_ld();
                    bootstrap_strip_i8(v, static_cast<uint32_t>(t) * kTileN + c * 32, c, n, st, mt, top);
                }
                LogState ls;
                ls.tau_l = ninf;
                ls.tau_g = st.tau_g;
                ls.thr = st.thr;
                log_boot_exchange(top, a.boot_j, a.boot_m, row_valid,
                                  a.ws_boot + static_cast<size_t>(row) * a.gpad, group, n_groups,
                                  a.ws_arrive + (row0 >> 5), ls, lane);
                st.tau_g = ls.tau_g;
                st.thr = ls.thr;
            }
            uint32_t g_next = __ldcg(wtau + lane);
            if (a.epi_mode == 0 || a.epi_mode == 3) {
                // two strips in registers: the TMEM load of strip c + 1 is in flight while strip c is
                // processed (the epilogue, not the tensor pipe, bounds this kernel)
                uint32_t va[32], vb[32];
                ptx::tmem_ld_32x32(taddr, va);
#pragma unroll 1
                for (int c = 0; c < kTileN / 32; c += 2) {
                    if ((c & 3) == 0) {
                        if (g_next > st.tau_g) {
                            st.tau_g = g_next;
                            st.thr = thr_of(ninf, st.tau_g);
                        }
                        g_next = __ldcg(wtau + lane);
                    }
                    ptx::tmem_wait_ld();
                    ptx::tmem_ld_32x32(taddr + (c + 1) * 32, vb);
                    process_strip_i8(va, static_cast<uint32_t>(t) * kTileN + c * 32, c, n, row_valid, st, mt, m_max, lg, lane);
                    ptx::tmem_wait_ld();
                    if (c + 2 < kTileN / 32) ptx::tmem_ld_32x32(taddr + (c + 2) * 32, va);
                    process_strip_i8(vb, static_cast<uint32_t>(t) * kTileN + (c + 1) * 32, c + 1, n, row_valid, st, mt, m_max, lg, lane);
                }
            } else if (a.epi_mode == 1) {                            // diagnostics: TMEM reads only
#pragma unroll 1
                for (int c = 0; c < kTileN / 32; ++c) {
                    uint32_t v[32];
                    ptx::tmem_ld_32x32(taddr + c * 32, v);
                    ptx::tmem_wait_ld();
                    asm volatile("" ::"r"(v[0]), "r"(v[31]));
                }
            }
            ptx::tc_fence_before();
            __syncwarp();
            if (lane == 0) {
                if constexpr (CG == 1) ptx::mbar_arrive(bar_tempty + 8 * acc);
                else ptx::mbar_arrive_cluster(bar_tempty + 8 * acc, 0);
            }
        }
        if (row_valid) a.ws_ccounts[static_cast<size_t>(row) * a.gpad + group] = st.cnt;   // read by the exact pass
        __syncwarp();
        if (lane == 0) atomicAdd(epi_done, 1u);
    }

    ptx::tc_fence_before();
    if constexpr (CG == 2) ptx::cluster_sync_all(); else __syncthreads();
    if (warp == 1) {
        ptx::tc_fence_after();
        ptx::tmem_dealloc<CG>(tmem_base, kTmemCols);
    }
}

// ------------------------------------------------------------------------------------------
// raw fp32 queries -> stored unit queries + int8 quantisation + query constants (one warp each)
// ------------------------------------------------------------------------------------------
template <typename T> __device__ __forceinline__ void store_query_elem(T* row, int i, float x);
template <> __device__ __forceinline__ void store_query_elem<float>(float* row, int i, float x) { row[i] = x; }
template <> __device__ __forceinline__ void store_query_elem<__nv_bfloat16>(__nv_bfloat16* row, int i, float x) {
    row[i] = __float2bfloat16_rn(x);                            // x is already a bf16 value: exact
}
template <> __device__ __forceinline__ void store_query_elem<__half>(__half* row, int i, float x) {
    row[i] = __float2half_rn(x);
}
template <> __device__ __forceinline__ void store_query_elem<Bf16x2>(Bf16x2* row, int i, float x) {
    __nv_bfloat16 hi, lo;
    split_bf16x2(x, hi, lo);                                    // x = hi + lo exactly
    row[i].v = hi;
    row[i + kDim].v = lo;
}

template <typename T>
__global__ void __launch_bounds__(128)
prepare_queries_kernel(const float* __restrict__ Q_raw, int b, T* __restrict__ Qst, int8_t* __restrict__ Q8,
                       float4* __restrict__ qmeta) {
    __shared__ __align__(16) float s_q[4][kDim];
    __shared__ __align__(16) float s_tile[4][8 * kNormBlockStride];
    const int lane = threadIdx.x & 31;
    const int warp = threadIdx.x >> 5;
    const int query = blockIdx.x * 4 + warp;
    if (query >= b) return;
    float* sq_ = s_q[warp];
    normalize_query_to_smem<T>(Q_raw + static_cast<int64_t>(query) * kDim, sq_, s_tile[warp], lane);
    __syncwarp();
    T* qrow = Qst + static_cast<int64_t>(query) * Elem<T>::kRowElems;
    float qv[32];
    float mx = 0.f, ss = 0.f;
#pragma unroll
    for (int j = 0; j < 32; ++j) {
        qv[j] = sq_[j * 32 + lane];
        store_query_elem<T>(qrow, j * 32 + lane, qv[j]);
        mx = fmaxf(mx, fabsf(qv[j]));
        ss = fmaf(qv[j], qv[j], ss);
    }
#pragma unroll
    for (int d = 16; d >= 1; d >>= 1) {
        mx = fmaxf(mx, __shfl_xor_sync(kFull, mx, d));
        ss += __shfl_xor_sync(kFull, ss, d);
    }
    const bool live = mx > 0.f && (mx - mx) == 0.f && (ss - ss) == 0.f;
    const float sq = live ? mx / 127.f : 0.f;
    const float inv = live ? 127.f / mx : 0.f;
    float e2 = 0.f;
    int8_t* q8 = Q8 + static_cast<int64_t>(query) * kDim;
#pragma unroll
    for (int j = 0; j < 32; ++j) {
        float r = rintf(qv[j] * inv);
        r = fminf(fmaxf(r, -127.f), 127.f);
        r = (r == r) ? r : 0.f;
        const float err = fmaf(-sq, r, qv[j]);
        e2 = fmaf(err, err, e2);
        q8[j * 32 + lane] = static_cast<int8_t>(static_cast<int>(r));
    }
#pragma unroll
    for (int d = 16; d >= 1; d >>= 1) e2 += __shfl_xor_sync(kFull, e2, d);
    if (lane == 0) {
        // |eq| and |q| as upper bounds (topk_prefilter.cu); the margin of a row is then
        //   m = (qe nd + qn eps + slack qn (nd + eps)) inflate = qe' nd + qn' eps
        // NaN / inf in the query make them NaN / inf -> every pair is logged -> exact scan
        const float qe = sqrtf(e2) * k2i::kInflate + 1e-12f;
        const float qn = sqrtf(ss) * k2i::kInflate;
        float4 m;
        m.x = sq;
        m.y = (qe + k2i::kSlack * qn) * k2i::kInflate * 1.000001f;
        m.z = (qn + k2i::kSlack * qn) * k2i::kInflate * 1.000001f;
        m.w = 0.f;
        qmeta[query] = m;
    }
}

// ------------------------------------------------------------------------------------------
// exact pass: one CTA per query.  tau' = the bound the threshold warps published last (a lower
// bound of the k-th best L, hence of the k-th best score).  All logged keys of the query (its
// group logs + its spill area) form one flat index space that the 256 threads walk together
// (every load of a step is independent: a few L2 round trips per 4096 keys, not one per group):
// a key whose upper bound U = L + 2 m (m recomputed from the row constants) stays below tau'
// cannot be in the result; the surviving rows are staged in shared memory and scored with K3's
// arithmetic, RIF rows in flight per warp.
// ------------------------------------------------------------------------------------------
template <typename T, int R>
__global__ void __launch_bounds__(256)
batched_rescore_kernel(const T* __restrict__ D, uint32_t n, const T* __restrict__ Qst, int b, int b_pad, int k,
                       const uint64_t* __restrict__ ws_logs, const uint32_t* __restrict__ ws_counts,
                       const uint32_t* __restrict__ ws_over, const uint32_t* __restrict__ ws_spill_cnt,
                       const uint64_t* __restrict__ ws_spill, uint32_t* __restrict__ ws_tau, int n_groups, int gpad,
                       int log_cap, const float4* __restrict__ meta, const float4* __restrict__ qmeta,
                       float* __restrict__ out_score, int64_t* __restrict__ out_idx, int64_t idx_offset,
                       uint32_t* __restrict__ out_rescored, int report_logged) {
    using E = Elem<T>;
    constexpr int LOADS = E::kLoads;
    constexpr int PER = E::kPer;
    constexpr int QG = E::kQGroups;
    constexpr int L = 32 * R;
    constexpr int kWarps = 8;
    constexpr int RIF = (LOADS >= 8) ? 2 : 4;                    // rows in flight per warp
    constexpr int kChunk = 4096;                                // keys examined per round = staged rows at most
    __shared__ uint64_t s_lists[kWarps][L];
    __shared__ uint32_t s_rows[kChunk];
    __shared__ uint32_t s_pref[k2::kMaxGroups + 2];             // exclusive prefix of the segment lengths
    __shared__ uint32_t s_n;
    const int lane = threadIdx.x & 31;
    const int warp = threadIdx.x >> 5;
    const int query = blockIdx.x;
    // header contract: the published bounds are zero again after the call (each CTA its own entry,
    // after reading it; CTA 0 the entries of the queries this launch does not have)
    if (blockIdx.x == 0)
        for (int i = b + threadIdx.x; i < k2::kQueriesPerLaunch; i += blockDim.x) ws_tau[i] = 0u;
    if (query >= b) return;
    const uint32_t tau_o = ws_tau[query];
    __syncthreads();
    if (threadIdx.x == 0) ws_tau[query] = 0u;

    float q[32];
    {
        const T* qp = Qst + static_cast<int64_t>(query) * E::kRowElems + lane * PER;
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

    // K3's arithmetic, operation for operation (topk_gemv.cu), RIF rows at a time
    auto score_rows = [&](const uint32_t (&rows)[RIF], int valid) {
        uint4 raw[RIF][LOADS];
#pragma unroll
        for (int j = 0; j < RIF; ++j) {
            if (j < valid) {
                const T* rp = D + static_cast<int64_t>(rows[j]) * E::kRowElems + lane * PER;
#pragma unroll
                for (int c = 0; c < LOADS; ++c) raw[j][c] = ldg_stream(rp + E::load_off(c));
            } else {
#pragma unroll
                for (int c = 0; c < LOADS; ++c) raw[j][c] = make_uint4(0, 0, 0, 0);
            }
        }
#pragma unroll
        for (int j = 0; j < RIF; ++j) {
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
            float s = a0 + a1;
#pragma unroll
            for (int d = 16; d >= 1; d >>= 1) s += __shfl_xor_sync(kFull, s, d);
            if (j < valid) {
                const uint64_t key = make_key(s, rows[j]);
                if (key > worst) {
                    list.insert(key, lane);
                    worst = list.worst();
                }
            }
        }
    };

    unsigned rescored = 0u;                                      // meaningful in warp 0
    if (ws_over[query] != 0u) {
        // log and spill area of this query overflowed: exact scan of every row (K3's loop)
        for (uint32_t base_row = warp * RIF; base_row < n; base_row += kWarps * RIF) {
            uint32_t rows[RIF];
            int valid = 0;
#pragma unroll
            for (int j = 0; j < RIF; ++j) {
                rows[j] = base_row + j;
                valid += (base_row + j < n) ? 1 : 0;
            }
            score_rows(rows, valid);
        }
        rescored = n;
    } else {
        // segment g < n_groups: the log of group g; segment n_groups: the spill area
        const int n_seg = n_groups + 1;
        for (int g = threadIdx.x; g < n_seg; g += blockDim.x) {
            uint32_t c;
            if (g < n_groups) {
                c = __ldcg(ws_counts + static_cast<size_t>(query) * gpad + g);
                c = c > static_cast<uint32_t>(log_cap) ? static_cast<uint32_t>(log_cap) : c;
            } else {
                c = __ldcg(ws_spill_cnt + query);
                c = c > static_cast<uint32_t>(k2i::kSpillCap) ? static_cast<uint32_t>(k2i::kSpillCap) : c;
            }
            s_pref[g + 1] = c;
        }
        __syncthreads();
        if (threadIdx.x == 0) {
            s_pref[0] = 0u;
            for (int g = 0; g < n_seg; ++g) s_pref[g + 1] += s_pref[g];
        }
        __syncthreads();
        const uint32_t total = s_pref[n_seg];
        if (report_logged) rescored = total;                     // diagnostics: pairs logged instead of rows scored
        const float tau = tau_o ? from_orderable_u32(tau_o) : __int_as_float(0xff800000);
        const float4 qm = qmeta[query];
        for (uint32_t c0 = 0; c0 < total; c0 += kChunk) {
            if (threadIdx.x == 0) s_n = 0u;
            __syncthreads();
            const uint32_t c1 = min(total, c0 + static_cast<uint32_t>(kChunk));
            for (uint32_t idx = c0 + threadIdx.x; idx < c1; idx += blockDim.x) {
                int lo = 0, hi = n_seg;                           // the segment holding flat index idx
                while (hi - lo > 1) {
                    const int mid = (lo + hi) >> 1;
                    if (s_pref[mid] <= idx) lo = mid; else hi = mid;
                }
                const uint32_t off = idx - s_pref[lo];
                const uint64_t* lp = lo < n_groups ? ws_logs + (static_cast<size_t>(lo) * b_pad + query) * log_cap
                                                   : ws_spill + static_cast<size_t>(query) * k2i::kSpillCap;
                const uint64_t key = __ldcg(lp + off);
                const uint32_t row = key_row(key);
                if (row < n) {
                    const float4 m4 = __ldg(meta + row);
                    const float m = fmaf(qm.y, m4.z, qm.z * m4.y) + 1e-30f;
                    // U = s8 + m with s8 = L + m up to the rounding of the subtraction that made L:
                    // 1e-6 of slack covers it.  Negated comparison: NaN bounds are kept.
                    const float U = key_score(key) + 2.f * m + 1e-6f;
                    if (!(U < tau)) s_rows[atomicAdd(&s_n, 1u)] = row;
                }
            }
            __syncthreads();
            const uint32_t cand = s_n;
            for (uint32_t e0 = warp * RIF; e0 < cand; e0 += kWarps * RIF) {
                uint32_t rows[RIF];
                int valid = 0;
#pragma unroll
                for (int j = 0; j < RIF; ++j) {
                    rows[j] = (e0 + j < cand) ? s_rows[e0 + j] : 0u;
                    valid += (e0 + j < cand) ? 1 : 0;
                }
                score_rows(rows, valid);
            }
            if (!report_logged) rescored += cand;
            __syncthreads();
        }
    }

    // ---- 8 warp lists -> the query's result ----
    list.store(s_lists[warp], lane);
    __syncthreads();
    if (warp != 0) return;
#pragma unroll 1
    for (int w = 1; w < kWarps; ++w) {
        WarpList<R> other;
        other.load(s_lists[w], lane);
        list.merge_sorted(other.key, lane);
    }
    emit_topk<R>(list, k, lane, out_score + static_cast<int64_t>(query) * k,
                 out_idx + static_cast<int64_t>(query) * k, idx_offset);
    if (lane == 0 && out_rescored) out_rescored[query] = rescored;
}

// ------------------------------------------------------------------------------ host
static inline int r_for_k_i8(int k) { return k <= 32 ? 1 : k <= 64 ? 2 : 4; }
static constexpr int64_t kI8HdrBytes = 8192;      // [bounds 4 KB][arrival counters 4 KB]

static int64_t i8_stage_bytes(int dtype) {
    const int64_t row = (dtype == 0 || dtype == 3) ? 4096 : 2048;
    return static_cast<int64_t>(k2::kQueriesPerLaunch) * (row + kDim + 16);
}

// [bounds | arrivals][bound logs: groups x b_pad x 64 R keys][counts, boot, candidate counts: b_pad x
// gpad each][over, spill counts: b_pad each] -- all of that is zeroed per launch -- then, never
// zeroed: [candidate logs: groups x b_pad x 1024 keys][spill: b_pad x kSpillCap keys] and the query
// staging area (stored queries, int8 queries, query constants).
static int64_t i8_words_bytes(int64_t gpad) { return (3 * gpad + 2) * k2::kQueriesPerLaunch * 4; }
static int64_t i8_clog_offset(int sm_count) {
    const int64_t gpad = (sm_count + 31) & ~31;
    const int64_t blogs = static_cast<int64_t>(sm_count) * k2::kRowsPerCta * 256 * 8;      // 64 R <= 256
    return (kI8HdrBytes + blogs + i8_words_bytes(gpad) + 255) & ~static_cast<int64_t>(255);
}
static int64_t i8_clog_bytes(int sm_count) { return static_cast<int64_t>(sm_count) * k2::kRowsPerCta * k2i::kLogCapMax * 8; }
static int64_t i8_spill_bytes() { return static_cast<int64_t>(k2::kQueriesPerLaunch) * k2i::kSpillCap * 8; }

int64_t batched_i8_workspace_bytes(int64_t /*n*/, int /*b*/, int /*k*/, int dtype, int sm_count) {
    return i8_clog_offset(sm_count) + i8_clog_bytes(sm_count) + i8_spill_bytes() + i8_stage_bytes(dtype);
}

template <int R, int CG, bool DEEP>
static int launch_i8_k(const CUtensorMap& tq, const CUtensorMap& td, const K2I8Args& a, cudaStream_t stream) {
    using C = k2i::CfgI8<CG, R, DEEP>;
    cudaError_t e = cudaFuncSetAttribute(topk_batched_i8_kernel<R, CG, DEEP>,
                                         cudaFuncAttributeMaxDynamicSharedMemorySize, C::kSmemBytes);
    if (e != cudaSuccess) { set_error("search_batched_prefiltered: smem attribute: %s", cudaGetErrorString(e)); return -2; }
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
    e = cudaLaunchKernelEx(&cfg, topk_batched_i8_kernel<R, CG, DEEP>, tq, td, a);
    if (e != cudaSuccess) { set_error("search_batched_prefiltered: launch: %s", cudaGetErrorString(e)); return -2; }
    return 0;
}

template <typename T, int R>
static int launch_i8_t(const void* D, int dtype, int64_t n, const void* D8, const void* meta, const float* Q_raw,
                       int b, int k, float* out_score, int64_t* out_idx, int64_t idx_offset,
                       uint32_t* out_rescored, void* ws, int sm_count, cudaStream_t stream) {
    char* w = static_cast<char*>(ws);
    char* clogs = w + i8_clog_offset(sm_count);
    char* spill = clogs + i8_clog_bytes(sm_count);
    char* stage = spill + i8_spill_bytes();
    const int64_t row_bytes = static_cast<int64_t>(Elem<T>::kRowElems) * sizeof(T);
    T* Qst = reinterpret_cast<T*>(stage);
    int8_t* Q8 = reinterpret_cast<int8_t*>(stage + k2::kQueriesPerLaunch * row_bytes);
    float4* qmeta = reinterpret_cast<float4*>(stage + k2::kQueriesPerLaunch * (row_bytes + kDim));
    const int n_dtiles = static_cast<int>((n + k2::kTileN - 1) / k2::kTileN);

    K2I8Args a = {};
    a.n = static_cast<uint32_t>(n);
    a.k = k;
    a.n_dtiles = n_dtiles;
    a.ws_tau = reinterpret_cast<uint32_t*>(w);
    a.ws_arrive = reinterpret_cast<uint32_t*>(w + 4096);
    a.ws_logs = reinterpret_cast<uint64_t*>(w + kI8HdrBytes);
    a.meta = static_cast<const float4*>(meta);
    a.qmeta = qmeta;
    a.epi_mode = g_k2_epilogue_mode.load();                    // read once: the exact pass below uses the same value
    a.ws_prog = reinterpret_cast<uint32_t*>(w + 4096 + 2048);     // second half of the arrivals page (zeroed per launch)
    const int window_knob = g_k2_window.load();
    a.window = window_knob > 0 ? window_knob : (window_knob < 0 ? 0 : 8);
    a.ws_spill = reinterpret_cast<uint64_t*>(spill);
    a.ws_clogs = reinterpret_cast<uint64_t*>(clogs);
    a.log_cap = k2i::kLogCapMax;
    a.blog_cap = 64 * R;
    for (int q0 = 0; q0 < b; q0 += k2::kQueriesPerLaunch) {
        const int bc = (b - q0 < k2::kQueriesPerLaunch) ? (b - q0) : k2::kQueriesPerLaunch;
        const int cg = (bc > k2::kRowsPerCta) ? 2 : 1;
        const int q_tile_rows = k2::kRowsPerCta * cg;
        const int n_qt = (bc + q_tile_rows - 1) / q_tile_rows;
        int n_groups = (sm_count / cg) / n_qt;
        if (n_groups < 1) n_groups = 1;
        if (n_dtiles > 0 && n_groups > n_dtiles) n_groups = n_dtiles;
        if (n_groups > k2::kMaxGroups) n_groups = k2::kMaxGroups;
        const int64_t b_pad = static_cast<int64_t>(n_qt) * q_tile_rows;
        const int64_t logs = static_cast<int64_t>(n_groups) * b_pad * a.blog_cap * 8;        // bound logs (zeroed)
        a.b = bc;
        a.n_qt = n_qt;
        a.n_groups = n_groups;
        a.gpad = (n_groups + 31) & ~31;
        a.ws_counts = reinterpret_cast<uint32_t*>(w + kI8HdrBytes + logs);
        a.ws_boot = a.ws_counts + b_pad * a.gpad;
        a.ws_ccounts = a.ws_boot + b_pad * a.gpad;
        a.ws_over = a.ws_ccounts + b_pad * a.gpad;
        a.ws_spill_cnt = a.ws_over + b_pad;
        const int64_t used = kI8HdrBytes + logs + (3 * a.gpad + 2) * b_pad * 4;
        // instruction descriptor (kind::i8): D s32 [4,6) = 2, A / B signed 8-bit [7,10) = [10,13) = 1,
        // K-major both, N >> 3 [17,23), M >> 4 [24,29)
        a.idesc = (2u << 4) | (1u << 7) | (1u << 10) | (static_cast<uint32_t>(k2::kTileN >> 3) << 17) |
                  (static_cast<uint32_t>(q_tile_rows >> 4) << 24);
        a.boot_j = a.boot_m = 0;
        {
            int j = (k + 15) / 16;
            while (j <= 16 && (k + j - 1) / j > n_groups) ++j;
            if (j <= 16) {
                a.boot_j = j;
                a.boot_m = (k + j - 1) / j;
            }
        }
        cudaError_t e = cudaMemsetAsync(ws, 0, static_cast<size_t>(used), stream);
        if (e != cudaSuccess) { set_error("search_batched_prefiltered: memset: %s", cudaGetErrorString(e)); return -2; }
        prepare_queries_kernel<T><<<(bc + 3) / 4, 128, 0, stream>>>(Q_raw + static_cast<int64_t>(q0) * kDim, bc, Qst, Q8,
                                                                    qmeta);
        e = cudaGetLastError();
        if (e != cudaSuccess) { set_error("search_batched_prefiltered: prepare launch: %s", cudaGetErrorString(e)); return -2; }
        if (n_dtiles > 0) {
            CUtensorMap tq, td;
            int rc = make_tile_map(&tq, Q8, kTileMapInt8, static_cast<uint64_t>(bc), k2::kRowsPerCta);
            if (rc != 0) return rc;
            rc = make_tile_map(&td, D8, kTileMapInt8, static_cast<uint64_t>(n), k2::kTileN / cg);
            if (rc != 0) return rc;
            const bool deep = (cg == 2) && (n_qt == 1);
            if (cg == 1) rc = launch_i8_k<R, 1, false>(tq, td, a, stream);
            else if (deep) rc = launch_i8_k<R, 2, true>(tq, td, a, stream);
            else rc = launch_i8_k<R, 2, false>(tq, td, a, stream);
            if (rc != 0) return rc;
        }
        batched_rescore_kernel<T, R><<<bc, 256, 0, stream>>>(
            static_cast<const T*>(D), static_cast<uint32_t>(n), Qst, bc, static_cast<int>(b_pad), k, a.ws_clogs, a.ws_ccounts,
            a.ws_over, a.ws_spill_cnt, a.ws_spill, a.ws_tau, n_dtiles > 0 ? n_groups : 0, a.gpad, a.log_cap, a.meta,
            qmeta, out_score + static_cast<int64_t>(q0) * k, out_idx + static_cast<int64_t>(q0) * k, idx_offset,
            out_rescored ? out_rescored + q0 : nullptr, a.epi_mode == 3 ? 1 : 0);
        e = cudaGetLastError();
        if (e != cudaSuccess) { set_error("search_batched_prefiltered: rescore launch: %s", cudaGetErrorString(e)); return -2; }
    }
    return 0;
}

template <typename T>
static int launch_i8_r(const void* D, int dtype, int64_t n, const void* D8, const void* meta, const float* Q_raw, int b,
                       int k, float* out_score, int64_t* out_idx, int64_t idx_offset, uint32_t* out_rescored,
                       void* ws, int sm_count, cudaStream_t stream) {
    switch (r_for_k_i8(k)) {
        case 1: return launch_i8_t<T, 1>(D, dtype, n, D8, meta, Q_raw, b, k, out_score, out_idx, idx_offset, out_rescored, ws, sm_count, stream);
        case 2: return launch_i8_t<T, 2>(D, dtype, n, D8, meta, Q_raw, b, k, out_score, out_idx, idx_offset, out_rescored, ws, sm_count, stream);
        default: return launch_i8_t<T, 4>(D, dtype, n, D8, meta, Q_raw, b, k, out_score, out_idx, idx_offset, out_rescored, ws, sm_count, stream);
    }
}

int launch_search_batched_prefiltered(const void* D, int dtype, int64_t n, const void* D8, const void* meta,
                                      const float* Q_raw, int b, int k, float* out_score, int64_t* out_idx,
                                      int64_t idx_offset, uint32_t* out_rescored, void* ws, int64_t ws_bytes,
                                      int sm_count, cudaStream_t stream) {
    if (ws_bytes < batched_i8_workspace_bytes(n, b, k, dtype, sm_count)) {
        set_error("search_batched_prefiltered: workspace %lld < %lld bytes", (long long)ws_bytes,
                  (long long)batched_i8_workspace_bytes(n, b, k, dtype, sm_count));
        return -3;
    }
    switch (dtype) {
        case 0: return launch_i8_r<float>(D, dtype, n, D8, meta, Q_raw, b, k, out_score, out_idx, idx_offset, out_rescored, ws, sm_count, stream);
        case 1: return launch_i8_r<__nv_bfloat16>(D, dtype, n, D8, meta, Q_raw, b, k, out_score, out_idx, idx_offset, out_rescored, ws, sm_count, stream);
        case 2: return launch_i8_r<__half>(D, dtype, n, D8, meta, Q_raw, b, k, out_score, out_idx, idx_offset, out_rescored, ws, sm_count, stream);
        case 3: return launch_i8_r<Bf16x2>(D, dtype, n, D8, meta, Q_raw, b, k, out_score, out_idx, idx_offset, out_rescored, ws, sm_count, stream);
        default: set_error("search_batched_prefiltered: bad dtype %d", dtype); return -1;
    }
}

}  // namespace sqe
