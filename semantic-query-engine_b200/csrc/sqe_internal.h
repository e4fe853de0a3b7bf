// Internal launch interfaces between the C ABI (api.cu) and the kernels.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include <atomic>

namespace sqe {

// K1
int launch_normalize_cast(const float* in, void* out, int64_t n, int out_dtype, int sm_count,
                          cudaStream_t stream);

struct XchgArgs;             // sqe_select.cuh: fused exchange of the sharded mode (null = none)

// K3
int64_t gemv_workspace_bytes(int nq, int k, int sm_count);
int make_xchg_args(XchgArgs* x, int rank, int world, void* const* peer_buffers, int64_t cap, unsigned epoch,
                   int nq, int k);
int launch_topk_gemv(const void* D, int dtype, int64_t n, const void* Q, bool raw_q, int nq, int k,
                     float* out_score, int64_t* out_idx, int64_t idx_offset, void* ws,
                     int64_t ws_bytes, int sm_count, cudaStream_t stream, const XchgArgs* xchg = nullptr,
                     bool pdl = false);

// K3p  int8 prefilter + exact rescoring (topk_prefilter.cu)
int launch_quantize_rows(const void* D, int dtype, int64_t n, void* D8, void* meta, int sm_count,
                         cudaStream_t stream);
int64_t prefilter_workspace_bytes(int64_t n, int nq, int k, int sm_count);
int launch_topk_prefiltered(const void* D, int dtype, int64_t n, const void* D8, const void* meta,
                            const void* Q, bool raw_q, int nq, int k, float* out_score, int64_t* out_idx,
                            int64_t idx_offset, unsigned* out_stats, void* ws, int64_t ws_bytes,
                            int sm_count, cudaStream_t stream, const XchgArgs* xchg = nullptr, bool pdl = false);

// K2
int64_t batched_workspace_bytes(int64_t n, int b, int k, int sm_count);
int launch_topk_batched(const void* D, int dtype, int64_t n, const void* Q, int b, int k,
                        float* out_score, int64_t* out_idx, int64_t idx_offset, void* ws,
                        int64_t ws_bytes, int sm_count, cudaStream_t stream);

// K2p  int8 tensor-core prefilter + exact rescoring (topk_batched_i8.cu)
int64_t batched_i8_workspace_bytes(int64_t n, int b, int k, int dtype, int sm_count);
int launch_search_batched_prefiltered(const void* D, int dtype, int64_t n, const void* D8, const void* meta,
                                      const float* Q_raw, int b, int k, float* out_score, int64_t* out_idx,
                                      int64_t idx_offset, uint32_t* out_rescored, void* ws, int64_t ws_bytes,
                                      int sm_count, cudaStream_t stream);

// K4
int launch_merge_topk(const float* scores, const int64_t* idx, int lists, int b, int k_in,
                      int k_out, float* out_score, int64_t* out_idx, cudaStream_t stream);

// K4x fused exchange + merge over peer-mapped buffers
int64_t exchange_buffer_bytes(int world, int64_t cap);
int launch_exchange_merge(const float* scores, const int64_t* idx, int b, int k_in, int k_out,
                          int rank, int world, void* const* peer_buffers, int64_t cap,
                          unsigned epoch, unsigned wait_mask, float* out_score, int64_t* out_idx,
                          int sm_count, cudaStream_t stream);

// K5 epilogue: (score,idx)[b] -> (score, idx32, hit)
int launch_cache_finalize(const float* score, const int64_t* idx, int b, double threshold,
                          float* out_score, int32_t* out_idx, uint8_t* out_hit,
                          cudaStream_t stream);

// ENC  embedding encoder (encoder_gemm.cu, encoder_attn.cu, encoder_rows.cu)
int launch_encoder_gemm(const void* X, int64_t ldx, const void* W, const float* bias, int64_t m, int n, int k,
                        int epilogue, void* out0, int64_t ld0, void* out1, int64_t ld1, int n_split, int q_cols,
                        float q_scale, const float* residual, int64_t ldr, const float* res_stats,
                        const float* res_gamma, const float* res_beta, int sm_count, cudaStream_t stream);
int64_t encoder_gemm_small_workspace_bytes();
int launch_encoder_gemm_small(const void* X, int64_t ldx, const void* W, const float* bias, int64_t m, int n, int k,
                              int epilogue, void* out0, int64_t ld0, void* out1, int64_t ld1, int n_split, int q_cols,
                              float q_scale, const float* residual, int64_t ldr, const float* res_stats,
                              const float* res_gamma, const float* res_beta, void* workspace, int64_t workspace_bytes,
                              cudaStream_t stream);
int launch_encoder_attention(const void* qk, const void* vt, int64_t t_pad, const void* tiles, int n_tiles,
                             int max_len, void* ctx, cudaStream_t stream);
int launch_encoder_layernorm(const float* in, const float* gamma, const float* beta, float eps, int64_t rows,
                             float* out32, void* out16, float* stats, cudaStream_t stream);
int launch_encoder_embed_ln(const int32_t* ids, const int32_t* pos, const float* word, const float* position,
                            const float* type0, const float* gamma, const float* beta, float eps, int64_t rows,
                            int vocab, int max_pos, float* out32, void* out16, float* stats, cudaStream_t stream);
int launch_encoder_pool(const float* h, const int32_t* first_token, int n_seq, float* out, int64_t ldo,
                        const float* stats, const float* gamma, const float* beta, cudaStream_t stream);

void set_error(const char* fmt, ...);

// tuning knobs and diagnostics buffers (api.cu).  Process-wide, set from any thread while other threads
// launch: atomics, and a launcher reads a knob ONCE per call.
extern std::atomic<int> g_k2_cta_group;      // 0 auto, 1, 2
extern std::atomic<int> g_k2_epilogue_mode;  // 0 normal; diagnostics only: 1 = TMEM loads only, 2 = no epilogue (results invalid)
extern std::atomic<int> g_k2_d_hint;         // retired experiment (accepted, ignored)
extern std::atomic<int> g_k2_window;         // K2p sibling progress window: 0 = default (8 tiles), -1 = unbounded, n = n tiles
extern std::atomic<void*> g_enc_gemm_debug;  // encoder_gemm.cu: role timers per CTA, or null
extern std::atomic<void*> g_enc_attn_debug;  // encoder_attn.cu: phase time stamps per CTA, or null
extern std::atomic<int> g_enc_small;         // 0 = few-token forward passes take the swap-AB split-K GEMM, 1 = never
extern std::atomic<int> g_enc_gemm_form;     // 0 auto, 1 = 128 x 64 tiles, 2 = 256 x 256 tiles on CTA pairs (encoder_gemm.cu)
extern std::atomic<void*> g_k2_debug;        // device buffer [grid][8] u64 of role timers, or null

}  // namespace sqe
