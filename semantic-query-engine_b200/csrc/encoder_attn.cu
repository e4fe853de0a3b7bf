// Encoder self-attention, one shot per (sequence, head, 128-query tile) on tcgen05.
//
// BERT sequences are at most 512 tokens, so a whole row of scores fits the tensor memory:
// 128 queries (TMEM lanes) x up to 512 keys (fp32 columns) = all 512 columns of one SM.  No
// online softmax, no rescaling of partial outputs.  One CTA serves a range of query tiles of one
// (sequence, head): K [S x 64] and V^T [64 x S] are fetched once (TMA, 128-byte swizzle, K-major:
// the operand layout of the scoring kernels) and stay in shared memory; per 128-query tile:
//
//   1. TMA:  the Q tile [128 x 64] (the next one is fetched as soon as step 2 has read this one);
//   2. MMA:  S = Q K^T into TMEM (Q was scaled by 1/8 in the QKV epilogue);
//   3. eight warps (two per lane quarter; each pair of quarters-of-keys = one half of the 64-key
//            chunks): row maximum, then p = 2^((s - max) log2 e), masked beyond the sequence end,
//            rounded to fp16 and written IN the swizzled K-major operand layout into a ring of four
//            16 KB chunk buffers (two per half);
//   4. MMA:  O += P_chunk V_chunk as each chunk is complete (its buffer is released by the MMA's
//            commit), accumulator over score columns that are already consumed;
//   5. the same warps scale O by 1 / sum and store fp16 rows of the context matrix.
//
// Sequences are PACKED: the entry list gives (first token, length, first query row, query rows);
// keys beyond the end of a sequence are whatever follows in the packed buffer (the next sequence,
// or zero rows) and are masked by index.
//
// Roofline: tensor pipe nominally (4 * 128 * S * 64 FLOP per tile) but at head_dim 64 the MUFU
// (one ex2 per score, 16 per clock per SM) is the binding unit: 128 S / 16 cycles per tile, which is
// 4 * 64 * 16 = 4096 FLOP per clock per SM = half of the tensor rate.
#include "sqe_enc.cuh"

namespace sqe {
namespace enc {

constexpr int kAttnThreads = 288;              // warp 0: TMA + MMA; warps 1..8: softmax / output
constexpr int kQBytes = kBM * 128;             // 16 KB
constexpr int kKVBlockBytes = 64 * 128;        // 8 KB: 64 keys x 64 dims (K) or 64 dims x 64 keys (V^T)
constexpr int kPBlockBytes = kBM * 128;        // 16 KB: 128 queries x 64 keys
constexpr int kPSlots = 4;                     // two per half of the keys

struct AttnArgs {
    const int4* tiles;        // (first token of the sequence, its length, first query row, query rows)
    __half* ctx;              // [T, 1024]
    int s_max;                // keys staged per CTA: max sequence length rounded up to 64
    uint32_t tmem_cols;       // power of two >= max(s_max, 64)
    long long* dbg;           // diagnostics: [CTA][8] cycles per phase (null in production)
};

__device__ __forceinline__ float ex2_approx(float x) {
    float y;
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
    return y;
}

__host__ __device__ inline int attn_smem_bytes(int s_max) {
    return kQBytes + 2 * s_max * 128 + kPSlots * kPBlockBytes + 2 * 2 * kBM * 4 /* max, sum */ + 128 /* barriers */ +
           16 + 1024;
}

__global__ void __launch_bounds__(kAttnThreads, 1)
encoder_attention_kernel(const __grid_constant__ CUtensorMap tmap_qk, const __grid_constant__ CUtensorMap tmap_vt,
                         const AttnArgs a) {
    extern __shared__ uint8_t smem_raw[];
    const uint32_t raw_addr = ptx::smem_u32(smem_raw);
    const uint32_t base = (raw_addr + 1023u) & ~1023u;
    uint8_t* sm = smem_raw + (base - raw_addr);

    const int warp = __shfl_sync(kFull, static_cast<int>(threadIdx.x >> 5), 0);
    const int lane = threadIdx.x & 31;
    const int head = blockIdx.y;
    const int4 tile = __ldg(a.tiles + blockIdx.x);
    const int seq_start = tile.x, seq_len = tile.y, q_first = tile.z, q_rows = tile.w;
    const int n_qt = (q_rows + kBM - 1) / kBM;
    const int q_end = min(seq_len, q_first + q_rows);
    const int s_pad = (seq_len + 63) & ~63;                    // keys scored
    const int n_kc = s_pad >> 6;                               // chunks of 64 keys
    const int h0c = (n_kc + 1) >> 1;                           // chunks of the first half of the warps

    const int kv_bytes = a.s_max * 128;
    const uint32_t sm_q = base;
    const uint32_t sm_k = base + kQBytes;
    const uint32_t sm_vt = sm_k + kv_bytes;
    const uint32_t sm_p = sm_vt + kv_bytes;
    const int off_red = kQBytes + 2 * kv_bytes + kPSlots * kPBlockBytes;
    float* red_max = reinterpret_cast<float*>(sm + off_red);   // [2][128]
    float* red_sum = red_max + 2 * kBM;                        // [2][128]
    const uint32_t bar_base = base + off_red + 2 * 2 * kBM * 4;
    const uint32_t bar_kv = bar_base;           // K, V^T have landed
    const uint32_t bar_q = bar_base + 8;        // a Q tile has landed
    const uint32_t bar_s = bar_base + 16;       // scores complete (and Q, K read)
    const uint32_t bar_o = bar_base + 24;       // output complete
    const uint32_t bar_oe = bar_base + 32;      // output read (8 warp arrivals): TMEM may be overwritten
    const uint32_t bar_full = bar_base + 40;    // [4] a P chunk is written (4 warp arrivals)
    const uint32_t bar_free = bar_base + 72;    // [4] a P chunk buffer has been read by its MMAs
    uint32_t* tmem_ptr_smem = reinterpret_cast<uint32_t*>(sm + off_red + 2 * 2 * kBM * 4 + 128);

    const bool stamp = a.dbg != nullptr && threadIdx.x == 32;
    [[maybe_unused]] long long t_s = 0, t_p1 = 0, t_p2 = 0, t_o = 0, t_out = 0;
    const long long t_begin = stamp ? clock64() : 0;

    if (warp == 0) {
        if (lane == 0) {
            ptx::prefetch_tensormap(&tmap_qk);
            ptx::prefetch_tensormap(&tmap_vt);
            ptx::mbar_init(bar_kv, 1);
            ptx::mbar_init(bar_q, 1);
            ptx::mbar_init(bar_s, 1);
            ptx::mbar_init(bar_o, 1);
            ptx::mbar_init(bar_oe, 8);
            for (int i = 0; i < kPSlots; ++i) {
                ptx::mbar_init(bar_full + 8 * i, 4);
                ptx::mbar_init(bar_free + 8 * i, 1);
            }
            ptx::fence_barrier_init();
        }
        __syncwarp();
        ptx::tmem_alloc<1>(ptx::smem_u32(tmem_ptr_smem), a.tmem_cols);
    }
    ptx::tc_fence_before();
    __syncthreads();
    ptx::tc_fence_after();
    const uint32_t tmem_base = *tmem_ptr_smem;

    if (warp == 0) {
        if (lane == 0) {
            // ------------------------------------------------------- loads: K, V^T once; first Q tile
            ptx::mbar_expect_tx(bar_q, kQBytes);
            ptx::tma_load_2d(sm_q, &tmap_qk, head * kHeadDim, seq_start + q_first, bar_q);          // rows 0..63
            ptx::tma_load_2d(sm_q + kKVBlockBytes, &tmap_qk, head * kHeadDim, seq_start + q_first + 64, bar_q);
            ptx::mbar_expect_tx(bar_kv, 2 * n_kc * kKVBlockBytes);
            for (int kc = 0; kc < n_kc; ++kc)
                ptx::tma_load_2d(sm_k + kc * kKVBlockBytes, &tmap_qk, kHidden + head * kHeadDim,
                                 seq_start + 64 * kc, bar_kv);
            for (int kc = 0; kc < n_kc; ++kc)
                ptx::tma_load_2d(sm_vt + kc * kKVBlockBytes, &tmap_vt, seq_start + 64 * kc, head * kHeadDim, bar_kv);
            const uint64_t dq = make_sw128_desc(sm_q);
            constexpr uint32_t idesc_o = idesc_f16(kBM, kHeadDim);
            for (int it = 0; it < n_qt; ++it) {
                ptx::mbar_wait(bar_q, it & 1);
                if (it == 0) ptx::mbar_wait(bar_kv, 0);
                else ptx::mbar_wait(bar_oe, (it - 1) & 1);     // the previous tile's scores and output are consumed
                ptx::tc_fence_after();
                // --------------------------------------------------- S = Q K^T
                for (int n0 = 0; n0 < s_pad; n0 += 256) {
                    const int nn = (s_pad - n0 < 256) ? (s_pad - n0) : 256;
                    const uint32_t idesc = idesc_f16(kBM, nn);
                    const uint64_t dk = make_sw128_desc(sm_k + n0 * 128);
#pragma unroll
                    for (int k4 = 0; k4 < kHeadDim / kUmmaK; ++k4)
                        ptx::umma_f16<1>(tmem_base + n0, dq + 2 * k4, dk + 2 * k4, idesc, k4 != 0 ? 1u : 0u);
                }
                ptx::umma_commit(bar_s);
                if (it + 1 < n_qt) {                           // Q is free once the scores are complete
                    ptx::mbar_wait(bar_s, it & 1);
                    const int qr = seq_start + q_first + (it + 1) * kBM;
                    ptx::mbar_expect_tx(bar_q, kQBytes);
                    ptx::tma_load_2d(sm_q, &tmap_qk, head * kHeadDim, qr, bar_q);
                    ptx::tma_load_2d(sm_q + kKVBlockBytes, &tmap_qk, head * kHeadDim, qr + 64, bar_q);
                }
                // --------------------------------------------------- O += P_chunk V_chunk, chunks as they come
                bool first = true;
                for (int j = 0; j < h0c; ++j) {
#pragma unroll
                    for (int h = 0; h < 2; ++h) {
                        const int mine = h ? n_kc - h0c : h0c;          // chunks of this half
                        if (j >= mine) continue;
                        const int kc = h ? h0c + j : j;
                        const int u = it * mine + j;                    // this half's running chunk count
                        const int slot = 2 * h + (u & 1);
                        ptx::mbar_wait(bar_full + 8 * slot, (u >> 1) & 1);
                        ptx::tc_fence_after();
                        const uint64_t dp = make_sw128_desc(sm_p + slot * kPBlockBytes);
                        const uint64_t dv = make_sw128_desc(sm_vt + kc * kKVBlockBytes);
#pragma unroll
                        for (int k4 = 0; k4 < kChunkK / kUmmaK; ++k4)
                            ptx::umma_f16<1>(tmem_base, dp + 2 * k4, dv + 2 * k4, idesc_o, (first && k4 == 0) ? 0u : 1u);
                        first = false;
                        ptx::umma_commit(bar_free + 8 * slot);
                    }
                }
                ptx::umma_commit(bar_o);
            }
        }
    } else {
        // ---------------------------------------------------------------- softmax + output
        const int quarter = warp & 3;                          // TMEM lanes 32 q .. 32 q + 31
        const int half = (warp - 1) >> 2;                      // which half of the key chunks
        const int r = quarter * 32 + lane;                     // query row of the tile
        const int mine = half ? n_kc - h0c : h0c;              // chunks of this half
        const int col_base = half ? h0c * 64 : 0;
        const uint32_t taddr = tmem_base + (static_cast<uint32_t>(quarter * 32) << 16);
        constexpr float kLog2e = 1.4426950408889634f;
        const float ninf = __int_as_float(0xff800000);

        for (int it = 0; it < n_qt; ++it) {
            long long t0 = stamp ? clock64() : 0;
            ptx::mbar_wait(bar_s, it & 1);
            ptx::tc_fence_after();
            if (stamp) { const long long t1 = clock64(); t_s += t1 - t0; t0 = t1; }
            // pass 1: row maximum over this warp's keys.  Strips inside the sequence need no mask;
            // strips beyond its end are skipped (their P entries are zeros, written in pass 2).
            float m0 = ninf, m1 = ninf, m2 = ninf, m3 = ninf;
#pragma unroll 1
            for (int s = 0; s < 2 * mine; ++s) {
                const int c0 = col_base + 32 * s;
                if (c0 >= seq_len) break;
                uint32_t v[32];
                ptx::tmem_ld_32x32(taddr + c0, v);
                ptx::tmem_wait_ld();
                if (c0 + 32 <= seq_len) {
#pragma unroll
                    for (int j = 0; j < 32; j += 4) {
                        m0 = fmaxf(m0, __uint_as_float(v[j]));
                        m1 = fmaxf(m1, __uint_as_float(v[j + 1]));
                        m2 = fmaxf(m2, __uint_as_float(v[j + 2]));
                        m3 = fmaxf(m3, __uint_as_float(v[j + 3]));
                    }
                } else {
#pragma unroll
                    for (int j = 0; j < 32; ++j)
                        if (c0 + j < seq_len) m0 = fmaxf(m0, __uint_as_float(v[j]));
                }
            }
            red_max[half * kBM + r] = fmaxf(fmaxf(m0, m1), fmaxf(m2, m3));
            ptx::bar_sync_named(1, 256);
            const float mx = fmaxf(red_max[r], red_max[kBM + r]);  // finite: key 0 is always inside the sequence
            const float mb = mx * kLog2e;
            if (stamp) { const long long t1 = clock64(); t_p1 += t1 - t0; t0 = t1; }
            // pass 2: p = 2^(s log2 e - max log2 e), fp16, into the operand layout, chunk by chunk; the
            // row sum is taken in fp32 before the rounding (2^-12 relative, far below the fp16 output)
            float s0 = 0.0f, s1 = 0.0f;
#pragma unroll 1
            for (int j = 0; j < mine; ++j) {
                const int u = it * mine + j;
                const int slot = 2 * half + (u & 1);
                if (u >= 2) ptx::mbar_wait(bar_free + 8 * slot, ((u >> 1) - 1) & 1);   // its previous MMAs have read it
                const uint32_t blk = sm_p + slot * kPBlockBytes + r * 128;
#pragma unroll
                for (int hs = 0; hs < 2; ++hs) {
                    const int c0 = col_base + 64 * j + 32 * hs;
                    uint32_t h[16];
                    if (c0 >= seq_len) {
#pragma unroll
                        for (int e = 0; e < 16; ++e) h[e] = 0u;
                    } else {
                        uint32_t v[32];
                        ptx::tmem_ld_32x32(taddr + c0, v);
                        ptx::tmem_wait_ld();
                        if (c0 + 32 <= seq_len) {
#pragma unroll
                            for (int e = 0; e < 16; ++e) {
                                const float p0 = ex2_approx(fmaf(__uint_as_float(v[2 * e]), kLog2e, -mb));
                                const float p1 = ex2_approx(fmaf(__uint_as_float(v[2 * e + 1]), kLog2e, -mb));
                                s0 += p0;
                                s1 += p1;
                                const __half2 hh = __floats2half2_rn(p0, p1);
                                h[e] = *reinterpret_cast<const uint32_t*>(&hh);
                            }
                        } else {
#pragma unroll
                            for (int e = 0; e < 16; ++e) {
                                const float p0 = (c0 + 2 * e < seq_len) ? ex2_approx(fmaf(__uint_as_float(v[2 * e]), kLog2e, -mb)) : 0.0f;
                                const float p1 = (c0 + 2 * e + 1 < seq_len) ? ex2_approx(fmaf(__uint_as_float(v[2 * e + 1]), kLog2e, -mb)) : 0.0f;
                                s0 += p0;
                                s1 += p1;
                                const __half2 hh = __floats2half2_rn(p0, p1);
                                h[e] = *reinterpret_cast<const uint32_t*>(&hh);
                            }
                        }
                    }
                    // operand layout: row r at r * 128 B, 16-byte chunk c at (c ^ (r & 7))
#pragma unroll
                    for (int c = 0; c < 4; ++c) {
                        const uint32_t addr = blk + (static_cast<uint32_t>((4 * hs + c) ^ (r & 7)) << 4);
                        asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(addr), "r"(h[4 * c]),
                                     "r"(h[4 * c + 1]), "r"(h[4 * c + 2]), "r"(h[4 * c + 3])
                                     : "memory");
                    }
                }
                ptx::tc_fence_before();            // these score columns may now be overwritten by O
                ptx::fence_proxy_async_smem();     // P: generic-proxy stores -> tensor-core operand reads
                __syncwarp();
                if (lane == 0) ptx::mbar_arrive(bar_full + 8 * slot);
            }
            red_sum[half * kBM + r] = s0 + s1;
            ptx::bar_sync_named(1, 256);           // both halves' sums are in shared memory
            if (stamp) { const long long t1 = clock64(); t_p2 += t1 - t0; t0 = t1; }

            ptx::mbar_wait(bar_o, it & 1);
            ptx::tc_fence_after();
            if (stamp) { const long long t1 = clock64(); t_o += t1 - t0; t0 = t1; }
            const float inv = 1.0f / (red_sum[r] + red_sum[kBM + r]);
            {
                uint32_t v[32];
                ptx::tmem_ld_32x32(taddr + half * 32, v);      // this warp's 32 of the 64 output dims
                ptx::tmem_wait_ld();
                const int qrow = q_first + it * kBM + r;
                if (qrow < q_end) {
                    uint4* op = reinterpret_cast<uint4*>(a.ctx + static_cast<int64_t>(seq_start + qrow) * kHidden +
                                                         head * kHeadDim + half * 32);
#pragma unroll
                    for (int j = 0; j < 4; ++j) {
                        uint32_t w[4];
#pragma unroll
                        for (int e = 0; e < 4; ++e) {
                            const __half2 hh = __floats2half2_rn(__uint_as_float(v[8 * j + 2 * e]) * inv,
                                                                 __uint_as_float(v[8 * j + 2 * e + 1]) * inv);
                            w[e] = *reinterpret_cast<const uint32_t*>(&hh);
                        }
                        op[j] = make_uint4(w[0], w[1], w[2], w[3]);
                    }
                }
            }
            ptx::tc_fence_before();
            __syncwarp();
            if (lane == 0) ptx::mbar_arrive(bar_oe);
            if (stamp) t_out += clock64() - t0;
        }
    }

    ptx::tc_fence_before();
    __syncthreads();
    if (stamp) {
        long long* dbg = a.dbg + (static_cast<size_t>(blockIdx.y) * gridDim.x + blockIdx.x) * 8;
        dbg[0] = clock64() - t_begin;
        dbg[1] = t_s;
        dbg[2] = t_p1;
        dbg[3] = t_p2;
        dbg[4] = t_o;
        dbg[5] = t_out;
        dbg[6] = n_qt;
        unsigned long long gt;
        asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(gt));
        dbg[7] = static_cast<long long>(gt);
    }
    if (warp == 0) {
        ptx::tc_fence_after();
        ptx::tmem_dealloc<1>(tmem_base, a.tmem_cols);
    }
}

}  // namespace enc

std::atomic<void*> g_enc_attn_debug{nullptr};     // diagnostics: device buffer [entries * 16][8] i64 of cycles per phase

int launch_encoder_attention(const void* qk, const void* vt, int64_t t_pad, const void* tiles, int n_tiles,
                             int max_len, void* ctx, cudaStream_t stream) {
    using namespace enc;
    if (n_tiles == 0) return 0;
    AttnArgs a = {};
    a.tiles = static_cast<const int4*>(tiles);
    a.ctx = static_cast<__half*>(ctx);
    a.s_max = (max_len + 63) & ~63;
    a.dbg = static_cast<long long*>(g_enc_attn_debug.load());
    a.tmem_cols = 64;
    while (a.tmem_cols < static_cast<uint32_t>(a.s_max)) a.tmem_cols *= 2;
    CUtensorMap tq, tv;
    // Q | K: [t_pad, 2048], box = 64 elements x 64 rows (the Q tile is two boxes); V^T: [1024, t_pad]
    int rc = make_map_2d(&tq, qk, 2 * kHidden, static_cast<uint64_t>(t_pad), 2 * kHidden, 64);
    if (rc != 0) return rc;
    rc = make_map_2d(&tv, vt, static_cast<uint64_t>(t_pad), kHidden, static_cast<uint64_t>(t_pad), 64);
    if (rc != 0) return rc;
    const int smem = attn_smem_bytes(a.s_max);
    cudaError_t e = cudaFuncSetAttribute(encoder_attention_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    if (e != cudaSuccess) { set_error("encoder_attention: smem attribute: %s", cudaGetErrorString(e)); return -2; }
    encoder_attention_kernel<<<dim3(static_cast<unsigned>(n_tiles), kHeads), kAttnThreads, smem, stream>>>(tq, tv, a);
    e = cudaGetLastError();
    if (e != cudaSuccess) { set_error("encoder_attention: launch: %s", cudaGetErrorString(e)); return -2; }
    return 0;
}

}  // namespace sqe
