/*
 * sqe_b200.h -- C ABI of the B200-native retrieval hot path.
 *
 * This is the drop-in boundary for the similarity path of
 * NeuralRevenant/semantic-query-engine.  The reference has no FFI of its own
 * (it is two Python files); each entry point below replaces the numpy / external
 * service expression cited next to it (paths relative to the reference root).
 * A Python maintainer binds these with ctypes; the package
 * `semantic-query-engine_b200/` is exactly that binding plus the host mirror of
 * the reference's `OpenSearchIndexer` / `lfu_cache_get` / `lfu_cache_put`.
 *
 * Conventions
 *   - every pointer is a DEVICE pointer (HBM) unless its name ends in `_host`;
 *   - `stream` is a `cudaStream_t` passed as `void*` (NULL = legacy default stream);
 *   - all calls are asynchronous on `stream`; outputs are valid after the stream
 *     is synchronised;
 *   - return value: 0 = ok, <0 = error (see SQE_E_*); `sqe_last_error()` gives the
 *     text for the calling thread.  No exceptions cross this boundary;
 *   - rows are `dim` = 1024 elements (EMBED_DIM, app/main.py:38), row-major,
 *     16-byte aligned;
 *   - there is no CPU fallback: on a machine without an sm_100 device the compute
 *     entry points return SQE_E_CUDA.
 *
 * Ordering contract (everywhere): results are best-first by (score descending,
 * row index ascending) -- "first maximum wins", app/main.py:84, generalised to
 * k > 1 (equals `np.argsort(-s, kind="stable")[:k]`).  Empty slots (k > rows)
 * hold score = -inf, index = -1.
 */
#ifndef SQE_B200_H
#define SQE_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#if defined(__GNUC__)
#define SQE_API __attribute__((visibility("default")))
#else
#define SQE_API
#endif

#define SQE_ABI_VERSION 2
#define SQE_DIM 1024            /* app/main.py:38 EMBED_DIM */
#define SQE_MAX_K_GEMV 256      /* largest k of sqe_topk_gemv / sqe_merge_topk */
#define SQE_MAX_K_BATCHED 128   /* largest k of sqe_topk_batched */

/* storage types of a shard in HBM */
#define SQE_F32 0
#define SQE_BF16 1
#define SQE_F16 2
/* split bf16: every value stored as hi = bf16(x) and lo = bf16(x - hi); a row is
 * [hi[1024] | lo[1024]] = 4096 bytes; the stored value hi + lo is exact in fp32 (16 significant
 * bits).  Gives the fp32 tolerance class (1e-5) on the bf16 tensor cores: K2 accumulates
 * Ql.Dl + Qh.Dl + Ql.Dh + Qh.Dh (small terms first), K3 reconstructs hi + lo exactly. */
#define SQE_BF16X2 3

/* error codes */
#define SQE_OK 0
#define SQE_E_ARG (-1)          /* bad argument (dim, dtype, k, alignment, null) */
#define SQE_E_CUDA (-2)         /* CUDA runtime / driver error, or no sm_100 device */
#define SQE_E_WORKSPACE (-3)    /* workspace too small */
#define SQE_E_UNSUPPORTED (-4)  /* combination not implemented (see message) */

SQE_API int sqe_abi_version(void);
SQE_API const char *sqe_last_error(void);

/* Number of SMs / whether the current device is sm_100 (1) or not (0); <0 on error. */
SQE_API int sqe_device_info(int *sm_count, int *cc_major, int *cc_minor);

/*
 * K1  fused L2-normalise + cast (ingest and query side).
 * Replaces   norms = np.linalg.norm(E, axis=1, keepdims=True); E = E / (norms + 1e-9)
 *            app/main.py:315-316 (corpus), :353-354 (query), app/embedding_gen.py:215-216.
 * in  [n, dim] fp32;  out [n, dim] of `out_dtype` ([n, 2 dim] bf16 for SQE_BF16X2).  The fp32
 * result is bit-identical to the numpy expression (same pairwise summation order, IEEE sqrt and
 * divide); bf16 / fp16 are that value rounded to nearest even; SQE_BF16X2 is its hi/lo split.
 */
SQE_API int sqe_normalize_cast(const float *in, void *out, int64_t n, int dim, int out_dtype,
                       void *stream);

/*
 * K3  batch-1 (or few-query) exact cosine top-k: HBM-bound streaming GEMV.
 * Replaces the k-NN query app/main.py:356-367 (external HNSW) with exact scoring.
 * D [n, dim] stored unit rows (dtype), Q [nq, dim] stored unit queries (same dtype).
 * Each query streams the whole shard once.  out_score [nq, k] fp32, out_idx [nq, k] int64
 * (= idx_offset + local row).  1 <= k <= SQE_MAX_K_GEMV, n < 2^32 - 1.
 *
 * Workspace contract: the first 4096 bytes must be ZERO the first time a workspace is used
 * (cudaMemset once after allocation); the library leaves them zero after every call, so no
 * per-call memset is launched.
 *
 * sqe_search_gemv = K1 (query side) + K3 in ONE launch: Q_raw holds the RAW fp32 query
 * embeddings [nq, dim]; every CTA normalises its query itself with the arithmetic of
 * sqe_normalize_cast and rounds it to the shard's storage type.  Same results as
 * sqe_normalize_cast followed by sqe_topk_gemv, bit for bit.
 */
SQE_API int64_t sqe_topk_gemv_workspace_bytes(int nq, int k);
SQE_API int sqe_search_gemv(const void *D, int dtype, int64_t n, int dim, const float *Q_raw, int nq,
                    int k, float *out_score, int64_t *out_idx, int64_t idx_offset, void *workspace,
                    int64_t workspace_bytes, void *stream);
SQE_API int sqe_topk_gemv(const void *D, int dtype, int64_t n, int dim, const void *Q, int nq, int k,
                  float *out_score, int64_t *out_idx, int64_t idx_offset, void *workspace,
                  int64_t workspace_bytes, void *stream);

/*
 * K3p  the same exact top-k as sqe_topk_gemv at about half the HBM traffic: an int8 prefilter
 * with a rigorous error bound + exact rescoring of the rows the bound cannot rule out.
 * Results are BIT-IDENTICAL to sqe_topk_gemv (scores, rows, tie order) for every input; only
 * the number of bytes read differs (10.5 GB instead of 20.5 GB per query on a 10M x 1024 bf16
 * shard).  Replaces the same reference lines as K3 (app/main.py:356-367).
 *
 * sqe_quantize_rows (ingest side, after sqe_normalize_cast): stored rows D [n, dim] (dtype) ->
 *   D8   int8 [n, 1024]: rint(d / sd), sd = max|d| / 127 per row, and
 *   meta float [n, 4] = {sd, eps >= |d - sd d8|_2, nd >= |sd d8|_2, 0} (16-byte aligned).
 * sqe_topk_gemv_prefiltered: D, Q, k, outputs and idx_offset as in sqe_topk_gemv; D8 / meta as
 *   written by sqe_quantize_rows for the same rows; 1 <= nq <= SQE_MAX_NQ_PREFILTER;
 *   out_rescored (may be NULL) uint32 [nq]: how many rows went through the exact pass.
 *   Workspace: sqe_topk_gemv_prefiltered_workspace_bytes(n, nq, k) (it holds one fp32 upper
 *   bound per row and query); first 4096 bytes zero on first use, left zero by every call.
 */
#define SQE_MAX_NQ_PREFILTER 64
SQE_API int sqe_quantize_rows(const void *D, int dtype, int64_t n, int dim, void *D8, void *meta,
                      void *stream);
SQE_API int64_t sqe_topk_gemv_prefiltered_workspace_bytes(int64_t n, int nq, int k);
SQE_API int sqe_topk_gemv_prefiltered(const void *D, int dtype, int64_t n, int dim, const void *D8,
                              const void *meta, const void *Q, int nq, int k, float *out_score,
                              int64_t *out_idx, int64_t idx_offset, uint32_t *out_rescored,
                              void *workspace, int64_t workspace_bytes, void *stream);

/*
 * K2  batched exact cosine top-k on the tensor cores (tcgen05 + TMA), with the
 * per-tile top-k selection fused into the accumulator epilogue so the score matrix
 * never reaches HBM.  Same contract as sqe_topk_gemv; dtype must be SQE_BF16, SQE_F16 or
 * SQE_BF16X2 (fp32 shards: SQE_E_UNSUPPORTED, use sqe_topk_gemv); 1 <= k <= SQE_MAX_K_BATCHED;
 * any b >= 1 (128 queries per CTA, CTA pairs of 256 when b > 128; more than 1024 queries are
 * processed in several launches).  D and Q must be 16-byte aligned device pointers.  The
 * workspace needs no initialisation (it is cleared by the call) and its first 4096 bytes are
 * ZERO again when the call has finished, so the same workspace may afterwards be passed to
 * sqe_topk_gemv / sqe_search_gemv (sqe_cache_top1 relies on this).
 */
SQE_API int64_t sqe_topk_batched_workspace_bytes(int64_t n, int b, int k);
SQE_API int sqe_topk_batched(const void *D, int dtype, int64_t n, int dim, const void *Q, int b, int k,
                     float *out_score, int64_t *out_idx, int64_t idx_offset, void *workspace,
                     int64_t workspace_bytes, void *stream);

/*
 * Tuning knobs (process-wide, for tests and benchmarks; results never depend on them, except the
 * diagnostics-only epilogue mode).  sqe_tuning_set returns the previous value, or SQE_E_ARG for
 * an unknown knob / value.
 *   SQE_TUNE_K2_CTA_GROUP: 0 = choose (CTA pairs when more than 128 queries are in flight),
 *                          1 = single-CTA UMMA (M = 128), 2 = CTA-pair UMMA (cta_group::2, M = 256).
 *   SQE_TUNE_K2_EPILOGUE_MODE: DIAGNOSTICS ONLY, results are invalid unless 0.  1 = the epilogue
 *                          only reads TMEM, 2 = no epilogue (isolates the TMA + MMA main loop).
 *   SQE_TUNE_K2_D_HINT:    RETIRED round-1 experiment (L2 policy of the shard-row TMA loads): a negative
 *                          result (profiles/README.md), removed from the kernel; accepted and ignored.
 *   SQE_TUNE_K2_WINDOW:    the int8 kernel (K2p) only: how many d-tiles a unit may run ahead of the
 *                          slowest unit that shares its d-tiles.  0 = default (8), -1 = unbounded,
 *                          n > 0 = n tiles.  (For the bf16 kernel the window was a negative result in
 *                          round 1 and is not compiled in.)
 *   SQE_TUNE_ENC_GEMM_FORM: tile geometry of sqe_encoder_gemm: 0 = choose by the number of tiles,
 *                          1 = 128 x 64 tiles (single CTA), 2 = 256 x 256 tiles (CTA pairs), 3 = the same in
 *                          clusters of four CTAs that share the X tile by TMA multicast.
 *   The role timers and the epilogue mode exist only in the diagnostics instantiations of the two
 *   benchmarked kernel forms (k <= 32 CTA-pair form with shared d-tiles; k > 64 CTA-pair form with
 *   one q-tile); every other form ignores them.
 */
#define SQE_TUNE_K2_CTA_GROUP 0
#define SQE_TUNE_K2_EPILOGUE_MODE 1
#define SQE_TUNE_K2_D_HINT 2
#define SQE_TUNE_K2_WINDOW 3
#define SQE_TUNE_ENC_GEMM_FORM 4
#define SQE_TUNE_ENC_SMALL 5          /* 0 = few-token passes take sqe_encoder_gemm_small (default), 1 = never */
SQE_API int sqe_tuning_set(int knob, int value);

/*
 * Diagnostics: when `device_buffer` is non-NULL, every later sqe_topk_batched launch writes
 * per-CTA role timers (clock64 cycles) and counters into it as u64 [grid][40]:
 *   0 producer total, 1 producer waiting for a free stage, 2 MMA issuer total, 3 MMA issuer
 *   waiting for operands, 4 MMA issuer waiting for a drained accumulator; then for each of the
 *   four epilogue warps w at 8 + 6 w: total, waiting for an accumulator, inside list merges,
 *   strips on the slow path, lists merged, slow-path column branches.
 * NULL switches it off (the default).
 */
SQE_API void sqe_debug_k2_timers(void *device_buffer);
/* Diagnostics: phase time stamps of sqe_encoder_attention, int64 [n_tiles * 16][8] per launch
 * (clock64 at: start, setup done, scores ready, maxima done, P written, output ready, end; then the
 * global timer at the end); NULL = off (the default). */
SQE_API void sqe_debug_encoder_attention_timers(void *device_buffer);
/* Diagnostics: role timers of sqe_encoder_gemm, int64 [grid][8] per launch (cycles: producer total / waiting for
 * free stages; MMA issuer total / waiting for data / waiting for the epilogue; tiles; global timer); NULL = off. */
SQE_API void sqe_debug_encoder_gemm_timers(void *device_buffer);

/*
 * K5  query-cache lookup: top-1 + similarity threshold.
 * Replaces the scan in lfu_cache_get, app/main.py:73-90: running maximum with strict `>`
 * from -1.0 (first maximum = lowest row wins), miss iff best < threshold.
 * C [n, dim] stored unit cache rows, Q [b, dim] stored unit queries (same dtype).
 * out_idx [b] int32 (-1 when no row beats -1.0 or the cache is empty), out_score [b] fp32,
 * out_hit [b] uint8 (1 iff out_idx >= 0 and !((double)score < threshold)): the comparison is done
 * in double like the reference's Python floats (best_sim < CACHE_SIM_THRESHOLD, 0.96 is not an
 * fp32 number).
 * `path`: 0 = choose (tensor cores when dtype is 16-bit and b > 1), 1 = force GEMV, 2 = force
 * tensor cores.
 * Workspace: same contract as sqe_topk_gemv (first 4096 bytes zero on first use; every call, on
 * either path and for any b, leaves them zero), so one workspace serves all lookups of a cache.
 */
SQE_API int64_t sqe_cache_top1_workspace_bytes(int64_t n, int b);
SQE_API int sqe_cache_top1(const void *C, int dtype, int64_t n, int dim, const void *Q, int b,
                   double threshold, float *out_score, int32_t *out_idx, uint8_t *out_hit,
                   int path, void *workspace, int64_t workspace_bytes, void *stream);

/*
 * K4  merge of per-shard top-k lists (after the all-gather of the corpus-sharded mode).
 * scores/idx [lists, b, k_in] (best-first per list, idx already global, -1 = empty);
 * out [b, k_out].  Global indices must be < 2^32 - 1.  k_in, k_out <= SQE_MAX_K_GEMV.
 */
SQE_API int sqe_merge_topk(const float *scores, const int64_t *idx, int lists, int b, int k_in,
                   int k_out, float *out_score, int64_t *out_idx, void *stream);

/*
 * K4x  fused exchange + merge of the corpus-sharded mode: ONE kernel per rank pushes its
 * [b, k_in] lists into every rank's peer-mapped gather buffer over NVLink, publishes an epoch
 * flag, waits for all ranks' flags in its own buffer and merges the `world` lists per query
 * (replaces two all-gathers + sqe_merge_topk).
 *   peer_buffers_host : HOST array of `world` DEVICE pointers, entry g = rank g's buffer as
 *                       mapped into this process (symmetric memory / CUDA IPC); every buffer
 *                       is sqe_exchange_buffer_bytes(world, capacity_entries) bytes, zeroed once
 *                       before the first call (all ranks, followed by a barrier);
 *   capacity_entries  : >= b * k_in, the same on every rank;
 *   epoch             : 1, 2, 3, ... the same on every rank for the same call;
 *   wait_mask         : bit g = wait for rank g; normally (1 << world) - 1 (tests that replay the
 *                       ranks one after the other on one GPU pass 0 for all but the last).
 * world <= 16, k_in, k_out <= SQE_MAX_K_GEMV, global rows < 2^32 - 1.
 */
SQE_API int64_t sqe_exchange_buffer_bytes(int world, int64_t capacity_entries);
SQE_API int sqe_exchange_merge(const float *scores, const int64_t *idx, int b, int k_in, int k_out,
                       int rank, int world, void *const *peer_buffers_host,
                       int64_t capacity_entries, uint32_t epoch, uint32_t wait_mask,
                       float *out_score, int64_t *out_idx, void *stream);

/*
 * Sharded ONE- or TWO-query search in a single launch per rank (north_star subsystem 4 for the
 * batch-1 path): the fused normalise + scan of sqe_search_gemv, and in the query's last CTA --
 * the one that has just merged the rank-local top-k -- the K4x exchange: push the list into every
 * rank's gather buffer, flag, wait for all ranks, merge, write the GLOBAL top-k.  No exchange
 * kernel, no second launch.  Buffers, capacity and epochs exactly as for sqe_exchange_merge (the
 * two may be mixed on the same buffers: one epoch per call, the same sequence on every rank);
 * 1 <= nq <= SQE_MAX_NQ_FUSED_EXCHANGE; world <= 1 or peer_buffers_host == NULL: plain
 * sqe_search_gemv.  out_idx are global rows (idx_offset + local row), < 2^32 - 1.
 *
 * sqe_search_gemv_prefiltered: the same for the prefiltered scan (K3p) -- RAW fp32 queries, the
 * query normalisation fused into both passes (no K1 launch), the exchange fused into the
 * rescoring pass.  Results are bit-identical to sqe_normalize_cast + sqe_topk_gemv_prefiltered
 * (+ sqe_exchange_merge).  With world <= 1: 1 <= nq <= SQE_MAX_NQ_PREFILTER.
 *
 * flags: SQE_FLAG_QUERIES_READY -- the caller states that Q_raw was complete BEFORE the previous
 * kernel of this stream was launched (a resident query, or one that arrived through a copy the
 * stream waited for).  The scan is then launched with programmatic stream serialization: it
 * starts while the previous scan of the stream is still in its tail (last-CTA merge, exchange
 * with the other ranks) and waits for it only before touching shared state.  Without the flag the
 * launch is an ordinary one.  Results never depend on it.
 */
#define SQE_MAX_NQ_FUSED_EXCHANGE 2
#define SQE_FLAG_QUERIES_READY 1
SQE_API int sqe_search_gemv_sharded(const void *D, int dtype, int64_t n, int dim, const float *Q_raw,
                            int nq, int k, float *out_score, int64_t *out_idx, int64_t idx_offset,
                            int rank, int world, void *const *peer_buffers_host,
                            int64_t capacity_entries, uint32_t epoch, int flags, void *workspace,
                            int64_t workspace_bytes, void *stream);
SQE_API int sqe_search_gemv_prefiltered(const void *D, int dtype, int64_t n, int dim, const void *D8,
                                const void *meta, const float *Q_raw, int nq, int k,
                                float *out_score, int64_t *out_idx, int64_t idx_offset,
                                uint32_t *out_rescored, int rank, int world,
                                void *const *peer_buffers_host, int64_t capacity_entries,
                                uint32_t epoch, int flags, void *workspace, int64_t workspace_bytes,
                                void *stream);

/*
 * K2p  the batched counterpart of K3p: exact cosine top-k for a BATCH of queries from the int8
 * copy of the shard on the tensor cores (tcgen05.mma kind::i8: half the bytes of a 16-bit shard,
 * twice its tensor rate) + exact rescoring of the (query, row) pairs the rigorous bound cannot
 * rule out.  Results are BIT-IDENTICAL to sqe_normalize_cast + sqe_topk_gemv (K3) for every
 * input and every storage class of D -- fp32 shards included, which have no other tensor-core
 * path.  Replaces the same reference lines as K2 (app/main.py:356-367; with k = 1 the scan of
 * lfu_cache_get, app/main.py:73-90).
 *   D, dtype, n      the exact stored rows (any storage class), as for sqe_topk_gemv;
 *   D8, meta         their coarse copy as written by sqe_quantize_rows;
 *   Q_raw            RAW fp32 queries [b, dim] (normalised and quantised by the call);
 *   1 <= k <= SQE_MAX_K_BATCHED, any b >= 1; out_score / out_idx [b, k] as for sqe_topk_gemv;
 *   out_rescored     (may be NULL) uint32 [b]: rows that went through the exact pass per query.
 * Workspace: sqe_search_batched_prefiltered_workspace_bytes(n, b, k, dtype); no initialisation
 * needed, first 4096 bytes zero again afterwards (shared-workspace contract).
 */
SQE_API int64_t sqe_search_batched_prefiltered_workspace_bytes(int64_t n, int b, int k, int dtype);
SQE_API int sqe_search_batched_prefiltered(const void *D, int dtype, int64_t n, int dim, const void *D8,
                                   const void *meta, const float *Q_raw, int b, int k,
                                   float *out_score, int64_t *out_idx, int64_t idx_offset,
                                   uint32_t *out_rescored, void *workspace, int64_t workspace_bytes,
                                   void *stream);

/*
 * K5 through K2p: the query-cache lookup (top-1 + threshold, app/main.py:73-90) for a batch of RAW
 * fp32 queries from the int8 copy of the cache rows + exact rescoring.  Outputs and threshold rule
 * as for sqe_cache_top1; the best entry and its score are K3's bit for bit ("first maximum wins" =
 * the lowest row), for any storage class of C.  C8 / meta: sqe_quantize_rows of the same rows.
 */
SQE_API int64_t sqe_cache_top1_prefiltered_workspace_bytes(int64_t n, int b, int dtype);
SQE_API int sqe_cache_top1_prefiltered(const void *C, int dtype, int64_t n, int dim, const void *C8,
                               const void *meta, const float *Q_raw, int b, double threshold,
                               float *out_score, int32_t *out_idx, uint8_t *out_hit, void *workspace,
                               int64_t workspace_bytes, void *stream);

/*
 * ENC  the embedding encoder in front of the path.
 * Replaces the HTTP round trip to the Ollama server that produces every embedding the path
 * consumes:  ollama_embed_text / embed_texts_in_batches / embed_query, app/main.py:134-180 and
 * app/embedding_gen.py:143-190 (model "mxbai-embed-large": a BERT-large encoder -- 24 post-LN
 * layers, hidden 1024, 16 heads of 64, FFN 4096 with erf-GELU, learned absolute positions <= 512,
 * CLS pooling; the arithmetic lives in the external server, so it is restated from the published
 * architecture, oracle/bert_oracle.py).  The host side (tokeniser, layer loop) is
 * semantic-query-engine_b200/encoder.py; these are its five kernels.  Tokens of all sequences of
 * a batch are PACKED into one [T, 1024] activation matrix (each sequence starts at the next multiple
 * of 8 rows, nothing else separates them); buffers
 * that feed a tensor-core operand are fp16, the residual stream and LayerNorm are fp32.
 *
 *   (plain forms first; the statistics forms that sqe_encoder_forward uses are described below)
 *   sqe_encoder_embed_ln   rows of word_emb[ids] + pos_emb[pos] + type_emb[0] -> LayerNorm ->
 *                          out_f32 [rows, 1024] and its fp16 copy out_f16; ids < 0 = padding row
 *                          (written as zeros).
 *   sqe_encoder_layernorm  in [rows, 1024] fp32 -> LayerNorm -> out_f32 + out_f16.
 *   sqe_encoder_gemm       Y = X W^T + bias on tcgen05 (X [m, k] fp16 with ldx elements per row,
 *                          W [n, k] fp16 = a torch Linear weight), n % 256 == 0, k % 64 == 0, fused
 *                          epilogue:
 *       SQE_ENC_EPI_SPLIT   columns < n_split -> fp16 out0 [m, ld0] (columns < q_cols scaled by
 *                           q_scale first), columns >= n_split -> fp16 TRANSPOSED
 *                           out1[(col - n_split), row] with ld1 elements per row (V^T of the attention);
 *                           n_split = n: a plain fp16 linear layer;
 *       SQE_ENC_EPI_RES_F32 + residual [m, ldr] fp32 -> fp32 out0 [m, ld0] (pre-LayerNorm sum);
 *       SQE_ENC_EPI_GELU    gelu(.) (erf form) -> fp16 out0 [m, ld0].
 *   sqe_encoder_attention  softmax(Q K^T) V per (sequence, head); qk [t_pad, 2048] fp16 = Q (already
 *                          scaled by 1/8) | K, vt [1024, t_pad] fp16 = V^T (t_pad % 8 == 0);
 *                          tiles int32 [n_tiles, 4] = (first token of the sequence, its length,
 *                          first query row of this entry, its number of query rows): one CTA per
 *                          (entry, head) keeps K / V of the sequence in shared memory and walks the
 *                          entry's 128-query tiles; the first token of a sequence
 *                          must be a multiple of 8 (a TMA box of V^T starts on a 16-byte boundary);
 *                          max_len = longest sequence (1..512); ctx [t_pad, 1024] fp16.
 *   sqe_encoder_pool       out[s, :] = h[first_token[s], :] (CLS pooling), ldo elements per row.
 */
#define SQE_ENC_HIDDEN 1024
#define SQE_ENC_MAX_TOKENS 512
#define SQE_ENC_EPI_SPLIT 0
#define SQE_ENC_EPI_RES_F32 1
#define SQE_ENC_EPI_GELU 2
/*
 * The LayerNorm STATISTICS form (what sqe_encoder_forward uses): a LayerNorm's fp32 output is only ever read
 * as the residual of the next residual GEMM, so it is never stored.  `stats` (float [rows][2] = {mean, rstd}):
 *   sqe_encoder_layernorm  stats != NULL: also writes the row statistics; out_f32 may then be NULL;
 *   sqe_encoder_embed_ln   stats != NULL: out_f32 receives the PRE-LayerNorm sum, stats its statistics;
 *   sqe_encoder_gemm(_small), SQE_ENC_EPI_RES_F32, res_stats != NULL: the residual operand is
 *       LayerNorm(residual) = ((residual - mean) * rstd) * res_gamma + res_beta, recomputed with the LayerNorm
 *       kernel's own operations (bit-identical to the output it would have stored);
 *   sqe_encoder_pool       stats != NULL: h holds pre-LayerNorm sums, out = LayerNorm(h[first_token]).
 * 10 -> 6 KB of HBM traffic per row and LayerNorm; results bit-identical to the plain form.
 */
SQE_API int sqe_encoder_embed_ln(const int32_t *ids, const int32_t *pos, const float *word_emb, int vocab,
                         const float *pos_emb, int max_pos, const float *type_emb, const float *gamma,
                         const float *beta, float eps, int64_t rows, float *out_f32, void *out_f16,
                         float *stats, void *stream);
SQE_API int sqe_encoder_layernorm(const float *in, const float *gamma, const float *beta, float eps,
                          int64_t rows, float *out_f32, void *out_f16, float *stats, void *stream);
SQE_API int sqe_encoder_gemm(const void *X, int64_t ldx, const void *W, const float *bias, int64_t m, int n,
                     int k, int epilogue, void *out0, int64_t ld0, void *out1, int64_t ld1, int n_split,
                     int q_cols, float q_scale, const float *residual, int64_t ldr, const float *res_stats,
                     const float *res_gamma, const float *res_beta, void *stream);
/* The same products for m <= 128 token rows (one query, a few queries): operands swapped (weight
 * rows are the M operand, n_tok = m rounded up to 16 the N operand), K split over ~100 CTAs, partial
 * tiles summed in split order by the last CTA of a feature tile (deterministic), same epilogues.
 * n % 128 == 0, n <= 4096; workspace: sqe_encoder_gemm_small_workspace_bytes(), zero-initialised once
 * (its tickets return to zero after every launch).  SQE_E_UNSUPPORTED for shapes it does not take. */
SQE_API int64_t sqe_encoder_gemm_small_workspace_bytes(void);
SQE_API int sqe_encoder_gemm_small(const void *X, int64_t ldx, const void *W, const float *bias, int64_t m, int n,
                           int k, int epilogue, void *out0, int64_t ld0, void *out1, int64_t ld1, int n_split,
                           int q_cols, float q_scale, const float *residual, int64_t ldr, const float *res_stats,
                           const float *res_gamma, const float *res_beta, void *workspace,
                           int64_t workspace_bytes, void *stream);
SQE_API int sqe_encoder_attention(const void *qk, const void *vt, int64_t t_pad, const int32_t *tiles,
                          int n_tiles, int max_len, void *ctx, void *stream);
SQE_API int sqe_encoder_pool(const float *h, const int32_t *first_token, int n_seq, float *out, int64_t ldo,
                     const float *stats, const float *gamma, const float *beta, void *stream);

/*
 * The whole forward pass in ONE call (what `GpuEmbeddingEncoder` issues per packed batch): embed_ln,
 * then per layer  QKV gemm -> attention -> output gemm (+ residual) -> LayerNorm -> FFN gemm (gelu) ->
 * FFN gemm (+ residual) -> LayerNorm,  then CLS pooling -- 2 + 7 n_layers kernel launches on `stream`,
 * nothing else (capturable in a CUDA graph).  rows_used = token rows of the packed batch that are in use
 * (<= t_pad; 0 = all): with rows_used <= 32 and a workspace in the buffers the linear layers take
 * sqe_encoder_gemm_small.  The structs hold DEVICE pointers; the structs
 * themselves (and the `layers` array) live in HOST memory and are read during the call only.
 */
typedef struct SqeEncoderLayer {
    const void *wqkv;           /* [3072, 1024] fp16: query | key | value Linear weights */
    const float *bqkv;          /* [3072] */
    const void *wo;             /* [1024, 1024] fp16 attention.output.dense */
    const float *bo, *ln1_gamma, *ln1_beta;
    const void *w1;             /* [intermediate, 1024] fp16 intermediate.dense */
    const float *b1;
    const void *w2;             /* [1024, intermediate] fp16 output.dense */
    const float *b2, *ln2_gamma, *ln2_beta;
} SqeEncoderLayer;

typedef struct SqeEncoderWeights {
    int n_layers, vocab, max_pos, intermediate;
    float eps;
    const float *word_emb, *pos_emb, *type_emb, *emb_gamma, *emb_beta;
    const SqeEncoderLayer *layers;      /* host array [n_layers] */
} SqeEncoderWeights;

typedef struct SqeEncoderBuffers {      /* activations of one packed batch, t_pad rows (t_pad % 128 == 0) */
    int64_t t_pad;
    float *sum_a, *sum_b;       /* [t_pad, 1024] pre-LayerNorm sums (the residual stream), ping-pong */
    float *stats_a, *stats_b;   /* [t_pad, 2] {mean, rstd} of the rows of sum_a / sum_b */
    void *h16;                  /* [t_pad, 1024] fp16 LayerNorm output (tensor-core operand) */
    void *qk;                   /* [t_pad, 2048] fp16 */
    void *vt;                   /* [1024, t_pad] fp16 */
    void *ctx;                  /* [t_pad, 1024] fp16 */
    void *ffn;                  /* [t_pad, intermediate] fp16 */
    void *small_ws;             /* zero-initialised workspace of sqe_encoder_gemm_small, or NULL */
    int64_t small_ws_bytes;
} SqeEncoderBuffers;

SQE_API int sqe_encoder_forward(const SqeEncoderWeights *weights_host, const SqeEncoderBuffers *buffers_host,
                        const int32_t *ids, const int32_t *pos, const int32_t *tiles, int n_tiles,
                        int max_len, const int32_t *first_token, int n_seq, int64_t rows_used, float *out,
                        int64_t ldo, void *stream);

#ifdef __cplusplus
}
#endif
#endif /* SQE_B200_H */
