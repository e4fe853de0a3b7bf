// K4x  fused exchange + merge of the corpus-sharded mode (north_star subsystem 4).
//
// After the local scan every rank holds [b, k] (score, global row) lists.  Instead of two
// NCCL all-gathers followed by a merge launch, ONE kernel per rank
//   1. pushes its lists into slot `rank` of EVERY rank's gather buffer with 128-bit stores
//      over NVLink (the buffers are peer-mapped symmetric memory, same layout everywhere),
//   2. publishes an epoch flag in every peer's buffer (system-scope release) once all of its
//      CTAs have pushed (last-CTA ticket),
//   3. waits until the flags of all ranks in its own buffer have reached the epoch
//      (system-scope acquire), and
//   4. merges the `world` lists of every query with the composite-key order (one warp per
//      query) and writes the result.
// The payload is tiny (b*k*16 B per peer, 164 KB at b=1024, k=10), so the exchange is latency
// bound; fusing it removes two collective launches and one kernel boundary per step.
//
// Buffer layout (identical on every rank, `sqe_exchange_buffer_bytes`):
//   [0, 64)             uint32 flag[g] = last epoch rank g has pushed into THIS buffer
//   [128, 132)          local CTA ticket (last-CTA-done pattern)
//   [256, ...)          records [2 parity][world][cap] of 16 bytes {int64 row; float score; u32 0}
// Calls alternate parity (epoch & 1): a rank can run at most one call ahead of a peer (its
// merge of call e needs the peer's push of call e, which the peer issues after finishing its
// merge of call e-1), so a peer never overwrites a slot that is still being read.
//
// Nothing like this exists in the reference (single process); the exchange is the only
// cross-GPU step of the path.
#include "sqe_common.cuh"
#include "sqe_internal.h"
#include "sqe_select.cuh"

namespace sqe {

// kMaxWorld, the header layout, PeerBufs, XRecord and the system-scope accessors: sqe_select.cuh

template <int R>
__global__ void __launch_bounds__(128)
exchange_merge_kernel(const float* __restrict__ scores, const int64_t* __restrict__ idx, int b,
                      int k_in, int k_out, int rank, int world, PeerBufs peers, int64_t cap,
                      unsigned epoch, unsigned wait_mask, float* __restrict__ out_score,
                      int64_t* __restrict__ out_idx) {
    const int lane = threadIdx.x & 31;
    const int parity = epoch & 1u;
    const int64_t n_entries = static_cast<int64_t>(b) * k_in;
    const int64_t slot_off = kXchgHeader + ((static_cast<int64_t>(parity) * world + rank) * cap) * 16;

    // ---- 1. push my lists into slot `rank` of every rank's buffer
    for (int64_t i = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x; i < n_entries;
         i += static_cast<int64_t>(gridDim.x) * blockDim.x) {
        XRecord rec;
        rec.row = idx[i];
        rec.score = scores[i];
        rec.pad = 0u;
        const uint4 raw = *reinterpret_cast<const uint4*>(&rec);
        for (int g = 0; g < world; ++g)
            *reinterpret_cast<uint4*>(peers.p[g] + slot_off + i * 16) = raw;
    }
    // ---- 2. last CTA of this rank publishes the epoch to every peer
    __threadfence_system();
    __syncthreads();
    if (threadIdx.x == 0) {
        unsigned* ticket = reinterpret_cast<unsigned*>(peers.p[rank]) + kXchgTicketWord;   // local
        const unsigned t = atomicAdd(ticket, 1u);
        if (t == gridDim.x - 1) {
            *ticket = 0u;                                   // ready for the next call
            fence_acq_rel_sys();                            // every CTA's pushes (seen via the ticket) first
            for (int g = 0; g < world; ++g)
                st_relaxed_sys(reinterpret_cast<unsigned*>(peers.p[g]) + rank, epoch);
        }
    }
    // ---- 3. wait for every rank's push into MY buffer
    const unsigned* my_flags = reinterpret_cast<const unsigned*>(peers.p[rank]);
    if (lane < world && ((wait_mask >> lane) & 1u)) {
        const long long t0 = clock64();
        unsigned spins = 0;
        while (static_cast<int>(ld_relaxed_sys(my_flags + lane) - epoch) < 0) {
            // a peer may legitimately arrive late (it is a collective: shard loading, a slow host);
            // give up only after ~10 minutes, like a collective library's watchdog would
            if ((++spins & 0xfffu) == 0 && clock64() - t0 > (1LL << 40)) __trap();
        }
        fence_acq_rel_sys();                                // flag reads before the data reads
    }
    __syncwarp();                                           // the polling lanes order the rest of the warp

    // ---- 4. merge: one warp per query
    const int warps_per_cta = blockDim.x >> 5;
    const char* mine = peers.p[rank] + kXchgHeader + (static_cast<int64_t>(parity) * world * cap) * 16;
    for (int query = blockIdx.x * warps_per_cta + (threadIdx.x >> 5); query < b;
         query += gridDim.x * warps_per_cta) {
        WarpList<R> list;
        list.clear();
        for (int g = 0; g < world; ++g) {
            const XRecord* recs = reinterpret_cast<const XRecord*>(mine + (static_cast<int64_t>(g) * cap) * 16) +
                                  static_cast<int64_t>(query) * k_in;
            WarpList<R> other;
#pragma unroll
            for (int r = 0; r < R; ++r) {
                const int i = r * 32 + lane;
                uint64_t key = 0ull;
                if (i < k_in) {
                    const uint4 raw = __ldcg(reinterpret_cast<const uint4*>(recs + i));
                    const XRecord rec = *reinterpret_cast<const XRecord*>(&raw);
                    if (rec.row >= 0) key = make_key(rec.score, static_cast<uint32_t>(rec.row));
                }
                other.key[r] = key;
            }
            other.sort(lane);                              // producers emit best-first; be robust anyway
            list.merge_sorted(other.key, lane);
        }
        emit_topk<R>(list, k_out, lane, out_score + static_cast<int64_t>(query) * k_out,
                     out_idx + static_cast<int64_t>(query) * k_out, 0);
    }
}

int64_t exchange_buffer_bytes(int world, int64_t cap) {
    return kXchgHeader + 2 * static_cast<int64_t>(world) * cap * 16;
}

int launch_exchange_merge(const float* scores, const int64_t* idx, int b, int k_in, int k_out,
                          int rank, int world, void* const* peer_buffers, int64_t cap,
                          unsigned epoch, unsigned wait_mask, float* out_score, int64_t* out_idx,
                          int sm_count, cudaStream_t stream) {
    if (world < 1 || world > kMaxWorld) { set_error("exchange: world=%d not in [1,%d]", world, kMaxWorld); return -1; }
    PeerBufs peers;
    for (int g = 0; g < kMaxWorld; ++g) peers.p[g] = (g < world) ? static_cast<char*>(peer_buffers[g]) : nullptr;
    const int kmax = k_in > k_out ? k_in : k_out;
    const int R = kmax <= 32 ? 1 : kmax <= 64 ? 2 : kmax <= 128 ? 4 : 8;
    // Every CTA waits on flags (its own rank's included, which the LAST local CTA publishes), so
    // the whole grid must be co-resident: cap it with the occupancy of THIS instantiation (the
    // register-resident lists of R = 4 / 8 allow fewer CTAs per SM than R = 1).
    int per_sm = 0;
    cudaError_t oe;
    switch (R) {
        case 1: oe = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, exchange_merge_kernel<1>, 128, 0); break;
        case 2: oe = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, exchange_merge_kernel<2>, 128, 0); break;
        case 4: oe = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, exchange_merge_kernel<4>, 128, 0); break;
        default: oe = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, exchange_merge_kernel<8>, 128, 0); break;
    }
    if (oe != cudaSuccess || per_sm < 1) {
        set_error("exchange: occupancy query: %s", cudaGetErrorString(oe));
        return -2;
    }
    if (per_sm > 4) per_sm = 4;                            // headroom for whatever else shares the SMs
    int grid = (b + 3) / 4;
    const int max_grid = sm_count * per_sm;
    if (grid > max_grid) grid = max_grid;
    if (grid < 1) grid = 1;
    dim3 g(grid), blk(128);
    switch (R) {
        case 1: exchange_merge_kernel<1><<<g, blk, 0, stream>>>(scores, idx, b, k_in, k_out, rank, world, peers, cap, epoch, wait_mask, out_score, out_idx); break;
        case 2: exchange_merge_kernel<2><<<g, blk, 0, stream>>>(scores, idx, b, k_in, k_out, rank, world, peers, cap, epoch, wait_mask, out_score, out_idx); break;
        case 4: exchange_merge_kernel<4><<<g, blk, 0, stream>>>(scores, idx, b, k_in, k_out, rank, world, peers, cap, epoch, wait_mask, out_score, out_idx); break;
        default: exchange_merge_kernel<8><<<g, blk, 0, stream>>>(scores, idx, b, k_in, k_out, rank, world, peers, cap, epoch, wait_mask, out_score, out_idx); break;
    }
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) { set_error("exchange: launch: %s", cudaGetErrorString(e)); return -2; }
    return 0;
}

}  // namespace sqe
