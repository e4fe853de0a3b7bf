// Shared tail of the streaming scans (K3 topk_gemv_kernel, K3p coarse_scan_kernel / rescore_kernel):
//   * 8 warp lists -> one CTA list -> (last CTA of the query, atomic ticket) every CTA list -> the
//     query's top list: one launch per query batch, no second kernel;
//   * optionally, in the corpus-sharded mode, the exchange with the other ranks from that same
//     last CTA: push the list into every rank's peer-mapped gather buffer, flag, wait, merge, write
//     the GLOBAL result -- the separate exchange kernel and its launch disappear for one or two
//     queries (north_star subsystem 4; the buffer layout is exchange.cu's).
#pragma once
#include "sqe_common.cuh"

namespace sqe {

constexpr int kMaxWorld = 16;
constexpr int kXchgHeader = 256;
constexpr int kXchgTicketWord = 32;          // u32 index of the local CTA ticket in the header
constexpr int kXchgFusedQueries = 2;         // per-query flag rows [q][16] fit the header in front of the ticket

struct PeerBufs {
    char* p[kMaxWorld];
};

struct __align__(16) XRecord {
    long long row;
    float score;
    unsigned pad;
};

// Exchange geometry of one call, passed by value to the scan kernels.  world <= 1: no exchange.
struct XchgArgs {
    PeerBufs peers;           // every rank's gather buffer as mapped into this process
    long long cap;            // entries per (parity, rank) slot
    int rank, world;
    unsigned epoch;
};

// One system-scope fence orders ALL the pushes before ALL the flag stores, and one after the
// poll loop orders the flag reads before the data reads; the flag accesses themselves are
// relaxed.  (st.release.sys per peer = one full fence per peer: measured ~3 us each.)
__device__ __forceinline__ void st_relaxed_sys(unsigned* p, unsigned v) {
    asm volatile("st.relaxed.sys.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ unsigned ld_relaxed_sys(const unsigned* p) {
    unsigned v;
    asm volatile("ld.relaxed.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void fence_acq_rel_sys() {
    asm volatile("fence.acq_rel.sys;" ::: "memory");
}

// Programmatic dependent launch (back-to-back scans of a query stream).  A scan kernel launched
// with the programmatic-serialization attribute may START while the previous kernel of the
// stream is still in its tail (last-CTA merge, exchange with the other ranks): its scan phase
// only reads the shard and the query.  `pdl_launch_dependents` (after the scan loop) lets the NEXT
// kernel start; `pdl_wait` (before the first access to anything an earlier kernel writes or still
// reads: workspace lists and tickets, outputs, exchange buffers, and for the rescoring pass its
// inputs tau / U) blocks until every earlier kernel has completed and its memory is visible.
// Both are no-ops in a normally launched kernel.  The attribute is only set when the caller
// states that the queries were complete before the previous launch (SQE_FLAG_QUERIES_READY).
__device__ __forceinline__ void pdl_launch_dependents() {
    asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
}
__device__ __forceinline__ void pdl_wait() {
    asm volatile("griddepcontrol.wait;" ::: "memory");
}

// Fold `count` sorted lists that sit `stride` lists apart (first one at `first`) into `list`.
// kBatch lists are fetched before the first of them is merged, so the L2 latency is paid once per
// batch instead of once per list (296 CTA lists per query: 37 dependent round trips per warp
// became 5).
template <int R>
__device__ __forceinline__ void fold_lists(WarpList<R>& list, const uint64_t* all, int first, int stride,
                                           int count, int lane) {
    constexpr int L = 32 * R;
    constexpr int kBatch = (R >= 8) ? 1 : 8 / R;
    for (int c0 = first; c0 < count; c0 += stride * kBatch) {
        uint64_t buf[kBatch][R];
#pragma unroll
        for (int j = 0; j < kBatch; ++j) {
            const int c = c0 + j * stride;
#pragma unroll
            for (int r = 0; r < R; ++r)
                buf[j][r] = c < count ? __ldcg(all + static_cast<int64_t>(c) * L + r * 32 + lane) : 0ull;
        }
#pragma unroll
        for (int j = 0; j < kBatch; ++j)
            if (c0 + j * stride < count) list.merge_sorted(buf[j], lane);
    }
}

// 8 warp lists -> CTA list -> (last CTA of the query) all CTA lists.  Returns true in warp 0 of
// the last CTA, with `list` = the merged result.  The caller resets `*counter` afterwards.
template <int R, int WARPS>
__device__ __forceinline__ bool merge_cta_and_grid(WarpList<R>& list, uint64_t (*s_lists)[32 * R],
                                                   int* s_is_last, uint64_t* ws_lists,
                                                   unsigned* counter, int query, int cta, int nctas,
                                                   int warp, int lane) {
    constexpr int L = 32 * R;
    list.store(s_lists[warp], lane);
    __syncthreads();
    uint64_t* my_slot = ws_lists + (static_cast<int64_t>(query) * nctas + cta) * L;
    if (warp == 0) {
#pragma unroll 1
        for (int w = 1; w < WARPS; ++w) {
            WarpList<R> other;
            other.load(s_lists[w], lane);
            list.merge_sorted(other.key, lane);
        }
        list.store(my_slot, lane);
        __threadfence();
        __syncwarp();
        if (lane == 0) {
            const unsigned ticket = atomicAdd(counter, 1u);
            *s_is_last = (ticket == static_cast<unsigned>(nctas) - 1) ? 1 : 0;
        }
    }
    __syncthreads();
    if (!*s_is_last) return false;
    __threadfence();
    const uint64_t* all = ws_lists + static_cast<int64_t>(query) * nctas * L;
    list.clear();
    fold_lists<R>(list, all, warp, WARPS, nctas, lane);         // each warp folds every WARPS-th CTA list
    __syncthreads();                                            // s_lists reuse
    list.store(s_lists[warp], lane);
    __syncthreads();
    if (warp != 0) return false;
#pragma unroll 1
    for (int w = 1; w < WARPS; ++w) {
        WarpList<R> other;
        other.load(s_lists[w], lane);
        list.merge_sorted(other.key, lane);
    }
    return true;
}

// Called by ONE warp (warp 0 of the query's last CTA) with the rank-local result in `list`
// (rows are shard-local; `idx_offset` makes them global).  Without an exchange: write it.  With
// one: push the k best as 16-byte records into slot `rank` of every rank's buffer, one
// system-scope fence, epoch flag [query][rank] into every buffer, poll this rank's own flags,
// one acquire fence, merge the `world` lists (composite key order: score desc, global row asc)
// and write the merged result.  Parity double-buffering and epochs as in exchange.cu: both
// kernels may be used on the same buffers, one epoch per call.
template <int R>
__device__ __forceinline__ void finish_query(const WarpList<R>& list, int k, int query, int lane,
                                             float* out_score, int64_t* out_idx, int64_t idx_offset,
                                             const XchgArgs& x) {
    float* os = out_score + static_cast<int64_t>(query) * k;
    int64_t* oi = out_idx + static_cast<int64_t>(query) * k;
    if (x.world <= 1) {
        emit_topk<R>(list, k, lane, os, oi, idx_offset);
        return;
    }
    const int parity = x.epoch & 1u;
    const int64_t slot_off = kXchgHeader + ((static_cast<int64_t>(parity) * x.world + x.rank) * x.cap +
                                            static_cast<int64_t>(query) * k) * 16;
#pragma unroll
    for (int r = 0; r < R; ++r) {
        const int i = r * 32 + lane;
        if (i < k) {
            const uint64_t key = list.key[r];
            XRecord rec;
            rec.row = key ? idx_offset + static_cast<long long>(key_row(key)) : -1;
            rec.score = key ? key_score(key) : __int_as_float(0xff800000);
            rec.pad = 0u;
            const uint4 raw = *reinterpret_cast<const uint4*>(&rec);
            for (int g = 0; g < x.world; ++g)
                *reinterpret_cast<uint4*>(x.peers.p[g] + slot_off + i * 16) = raw;
        }
    }
    __threadfence_system();                                     // every lane's pushes ...
    __syncwarp();
    const unsigned flag_word = static_cast<unsigned>(query) * kMaxWorld;
    if (lane < x.world)                                         // ... before the flags
        st_relaxed_sys(reinterpret_cast<unsigned*>(x.peers.p[lane]) + flag_word + x.rank, x.epoch);
    const unsigned* my_flags = reinterpret_cast<const unsigned*>(x.peers.p[x.rank]) + flag_word;
    if (lane < x.world) {
        const long long t0 = clock64();
        unsigned spins = 0;
        while (static_cast<int>(ld_relaxed_sys(my_flags + lane) - x.epoch) < 0) {
            // a peer may legitimately arrive late (a collective); give up after ~10 minutes
            if ((++spins & 0xfffu) == 0 && clock64() - t0 > (1LL << 40)) __trap();
        }
        fence_acq_rel_sys();                                    // flag reads before the data reads
    }
    __syncwarp();
    const char* mine = x.peers.p[x.rank] + kXchgHeader + (static_cast<int64_t>(parity) * x.world * x.cap) * 16;
    WarpList<R> merged;
    merged.clear();
    for (int g = 0; g < x.world; ++g) {
        const XRecord* recs = reinterpret_cast<const XRecord*>(mine + (static_cast<int64_t>(g) * x.cap) * 16) +
                              static_cast<int64_t>(query) * k;
        uint64_t other[R];
#pragma unroll
        for (int r = 0; r < R; ++r) {
            const int i = r * 32 + lane;
            uint64_t key = 0ull;
            if (i < k) {
                const uint4 raw = __ldcg(reinterpret_cast<const uint4*>(recs + i));
                const XRecord rec = *reinterpret_cast<const XRecord*>(&raw);
                if (rec.row >= 0) key = make_key(rec.score, static_cast<uint32_t>(rec.row));
            }
            other[r] = key;
        }
        merged.merge_sorted(other, lane);                       // every rank pushes a sorted list
    }
    emit_topk<R>(merged, k, lane, os, oi, 0);
}

}  // namespace sqe
