// The C ABI (include/sqe_b200.h): argument checks, error text, dispatch.  No torch
// types, no exceptions across the boundary.
#include <cstdarg>
#include <cstdio>
#include <cstring>

#include "../../include/sqe_b200.h"
#include "sqe_internal.h"
#include "sqe_select.cuh"

namespace sqe {

static thread_local char g_err[512] = "";
std::atomic<int> g_k2_cta_group{0};
std::atomic<void*> g_k2_debug{nullptr};
std::atomic<int> g_k2_epilogue_mode{0};
std::atomic<int> g_k2_d_hint{0};
std::atomic<int> g_k2_window{0};
std::atomic<int> g_enc_small{0};

void set_error(const char* fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
}

struct DevInfo {
    int ok = 0;
    int sm_count = 0;
    int major = 0;
    int minor = 0;
};

// per-device cache (8 GPUs per node)
static int device_info(DevInfo* out) {
    static DevInfo cache[64];
    int dev = 0;
    cudaError_t e = cudaGetDevice(&dev);
    if (e != cudaSuccess) {
        set_error("cudaGetDevice: %s", cudaGetErrorString(e));
        return SQE_E_CUDA;
    }
    if (dev < 0 || dev >= 64) { set_error("device ordinal %d out of range", dev); return SQE_E_CUDA; }
    if (!cache[dev].ok) {
        DevInfo d;
        if ((e = cudaDeviceGetAttribute(&d.sm_count, cudaDevAttrMultiProcessorCount, dev)) != cudaSuccess ||
            (e = cudaDeviceGetAttribute(&d.major, cudaDevAttrComputeCapabilityMajor, dev)) != cudaSuccess ||
            (e = cudaDeviceGetAttribute(&d.minor, cudaDevAttrComputeCapabilityMinor, dev)) != cudaSuccess) {
            set_error("cudaDeviceGetAttribute: %s", cudaGetErrorString(e));
            return SQE_E_CUDA;
        }
        d.ok = 1;
        cache[dev] = d;
    }
    *out = cache[dev];
    if (out->major != 10) {
        set_error("device is sm_%d%d; this library contains sm_100a code only (no fallback)",
                  out->major, out->minor);
        return SQE_E_CUDA;
    }
    return SQE_OK;
}

static bool aligned16(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15u) == 0; }

static int check_common(const char* who, const void* D, int dtype, int64_t n, int dim, const void* Q,
                        int nq) {
    if (dim != SQE_DIM) { set_error("%s: dim must be %d (got %d)", who, SQE_DIM, dim); return SQE_E_ARG; }
    if (dtype < SQE_F32 || dtype > SQE_BF16X2) { set_error("%s: bad dtype %d", who, dtype); return SQE_E_ARG; }
    if (n < 0 || n >= 0xffffffffLL) { set_error("%s: n=%lld out of range", who, (long long)n); return SQE_E_ARG; }
    if (nq < 0) { set_error("%s: negative query count", who); return SQE_E_ARG; }
    if ((n > 0 && D == nullptr) || (nq > 0 && Q == nullptr)) { set_error("%s: null pointer", who); return SQE_E_ARG; }
    if (!aligned16(D) || !aligned16(Q)) { set_error("%s: pointers must be 16-byte aligned", who); return SQE_E_ARG; }
    return SQE_OK;
}

}  // namespace sqe

using namespace sqe;

extern "C" {

int sqe_abi_version(void) { return SQE_ABI_VERSION; }

const char* sqe_last_error(void) { return g_err; }

int sqe_tuning_set(int knob, int value) {
    if (knob == SQE_TUNE_K2_CTA_GROUP && value >= 0 && value <= 2) {
        return g_k2_cta_group.exchange(value);
    }
    if (knob == SQE_TUNE_K2_EPILOGUE_MODE && value >= 0 && value <= 3) {
        return g_k2_epilogue_mode.exchange(value);
    }
    if (knob == SQE_TUNE_K2_WINDOW && value >= -1 && value <= 1024) {
        return g_k2_window.exchange(value);
    }
    if (knob == SQE_TUNE_K2_D_HINT && value >= 0 && value <= 4) {
        return g_k2_d_hint.exchange(value);
    }
    if (knob == SQE_TUNE_ENC_SMALL && value >= 0 && value <= 1) {
        return g_enc_small.exchange(value);
    }
    if (knob == SQE_TUNE_ENC_GEMM_FORM && value >= 0 && value <= 4) {
        return g_enc_gemm_form.exchange(value);
    }
    set_error("tuning_set: unknown knob %d / value %d", knob, value);
    return SQE_E_ARG;
}

void sqe_debug_k2_timers(void* device_buffer) { g_k2_debug = device_buffer; }

void sqe_debug_encoder_attention_timers(void* device_buffer) { g_enc_attn_debug = device_buffer; }

void sqe_debug_encoder_gemm_timers(void* device_buffer) { g_enc_gemm_debug = device_buffer; }

int sqe_device_info(int* sm_count, int* cc_major, int* cc_minor) {
    DevInfo d;
    int rc = device_info(&d);
    if (sm_count) *sm_count = d.sm_count;
    if (cc_major) *cc_major = d.major;
    if (cc_minor) *cc_minor = d.minor;
    if (rc != SQE_OK) return d.ok ? 0 : rc;
    return 1;
}

int sqe_normalize_cast(const float* in, void* out, int64_t n, int dim, int out_dtype, void* stream) {
    if (dim != SQE_DIM) { set_error("normalize_cast: dim must be %d (got %d)", SQE_DIM, dim); return SQE_E_ARG; }
    if (out_dtype < SQE_F32 || out_dtype > SQE_BF16X2) { set_error("normalize_cast: bad dtype %d", out_dtype); return SQE_E_ARG; }
    if (n < 0) { set_error("normalize_cast: negative n"); return SQE_E_ARG; }
    if (n == 0) return SQE_OK;
    if (!in || !out || !aligned16(in) || !aligned16(out)) { set_error("normalize_cast: null or unaligned pointer"); return SQE_E_ARG; }
    DevInfo d;
    int rc = device_info(&d);
    if (rc != SQE_OK) return rc;
    rc = launch_normalize_cast(in, out, n, out_dtype, d.sm_count, static_cast<cudaStream_t>(stream));
    if (rc != 0) return SQE_E_ARG;
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) { set_error("normalize_cast: launch: %s", cudaGetErrorString(e)); return SQE_E_CUDA; }
    return SQE_OK;
}

int64_t sqe_topk_gemv_workspace_bytes(int nq, int k) {
    DevInfo d;
    if (device_info(&d) != SQE_OK) d.sm_count = 160;       // upper bound when queried off-device
    if (nq < 1) nq = 1;
    if (k < 1) k = 1;
    return gemv_workspace_bytes(nq, k, d.sm_count);
}

int sqe_topk_gemv(const void* D, int dtype, int64_t n, int dim, const void* Q, int nq, int k,
                  float* out_score, int64_t* out_idx, int64_t idx_offset, void* workspace,
                  int64_t workspace_bytes, void* stream) {
    int rc = check_common("topk_gemv", D, dtype, n, dim, Q, nq);
    if (rc != SQE_OK) return rc;
    if (k < 1 || k > SQE_MAX_K_GEMV) { set_error("topk_gemv: k=%d not in [1,%d]", k, SQE_MAX_K_GEMV); return SQE_E_ARG; }
    if (nq == 0) return SQE_OK;
    if (!out_score || !out_idx || !workspace) { set_error("topk_gemv: null output/workspace"); return SQE_E_ARG; }
    DevInfo d;
    rc = device_info(&d);
    if (rc != SQE_OK) return rc;
    return launch_topk_gemv(D, dtype, n, Q, false, nq, k, out_score, out_idx, idx_offset, workspace,
                            workspace_bytes, d.sm_count, static_cast<cudaStream_t>(stream));
}

int sqe_search_gemv(const void* D, int dtype, int64_t n, int dim, const float* Q_raw, int nq, int k,
                    float* out_score, int64_t* out_idx, int64_t idx_offset, void* workspace,
                    int64_t workspace_bytes, void* stream) {
    int rc = check_common("search_gemv", D, dtype, n, dim, Q_raw, nq);
    if (rc != SQE_OK) return rc;
    if (k < 1 || k > SQE_MAX_K_GEMV) { set_error("search_gemv: k=%d not in [1,%d]", k, SQE_MAX_K_GEMV); return SQE_E_ARG; }
    if (nq == 0) return SQE_OK;
    if (!out_score || !out_idx || !workspace) { set_error("search_gemv: null output/workspace"); return SQE_E_ARG; }
    DevInfo d;
    rc = device_info(&d);
    if (rc != SQE_OK) return rc;
    return launch_topk_gemv(D, dtype, n, Q_raw, true, nq, k, out_score, out_idx, idx_offset, workspace,
                            workspace_bytes, d.sm_count, static_cast<cudaStream_t>(stream));
}

int sqe_quantize_rows(const void* D, int dtype, int64_t n, int dim, void* D8, void* meta, void* stream) {
    if (dim != SQE_DIM) { set_error("quantize_rows: dim must be %d (got %d)", SQE_DIM, dim); return SQE_E_ARG; }
    if (dtype < SQE_F32 || dtype > SQE_BF16X2) { set_error("quantize_rows: bad dtype %d", dtype); return SQE_E_ARG; }
    if (n < 0 || n >= 0xffffffffLL) { set_error("quantize_rows: n=%lld out of range", (long long)n); return SQE_E_ARG; }
    if (n == 0) return SQE_OK;
    if (!D || !D8 || !meta || !aligned16(D) || !aligned16(D8) || !aligned16(meta)) {
        set_error("quantize_rows: null or unaligned pointer");
        return SQE_E_ARG;
    }
    DevInfo d;
    int rc = device_info(&d);
    if (rc != SQE_OK) return rc;
    rc = launch_quantize_rows(D, dtype, n, D8, meta, d.sm_count, static_cast<cudaStream_t>(stream));
    return rc == 0 ? SQE_OK : (rc == -1 ? SQE_E_ARG : SQE_E_CUDA);
}

int64_t sqe_topk_gemv_prefiltered_workspace_bytes(int64_t n, int nq, int k) {
    DevInfo d;
    if (device_info(&d) != SQE_OK) d.sm_count = 160;
    if (nq < 1) nq = 1;
    if (k < 1) k = 1;
    if (n < 0) n = 0;
    return prefilter_workspace_bytes(n, nq, k, d.sm_count);
}

int sqe_topk_gemv_prefiltered(const void* D, int dtype, int64_t n, int dim, const void* D8,
                              const void* meta, const void* Q, int nq, int k, float* out_score,
                              int64_t* out_idx, int64_t idx_offset, uint32_t* out_rescored,
                              void* workspace, int64_t workspace_bytes, void* stream) {
    int rc = check_common("topk_gemv_prefiltered", D, dtype, n, dim, Q, nq);
    if (rc != SQE_OK) return rc;
    if (k < 1 || k > SQE_MAX_K_GEMV) { set_error("topk_gemv_prefiltered: k=%d not in [1,%d]", k, SQE_MAX_K_GEMV); return SQE_E_ARG; }
    if (nq > SQE_MAX_NQ_PREFILTER) { set_error("topk_gemv_prefiltered: nq=%d > %d", nq, SQE_MAX_NQ_PREFILTER); return SQE_E_ARG; }
    if (nq == 0) return SQE_OK;
    if (!out_score || !out_idx || !workspace) { set_error("topk_gemv_prefiltered: null output/workspace"); return SQE_E_ARG; }
    if (n > 0 && (!D8 || !meta || !aligned16(D8) || !aligned16(meta))) {
        set_error("topk_gemv_prefiltered: null or unaligned coarse rows");
        return SQE_E_ARG;
    }
    DevInfo d;
    rc = device_info(&d);
    if (rc != SQE_OK) return rc;
    rc = launch_topk_prefiltered(D, dtype, n, D8, meta, Q, false, nq, k, out_score, out_idx, idx_offset,
                                 out_rescored, workspace, workspace_bytes, d.sm_count,
                                 static_cast<cudaStream_t>(stream));
    return rc == 0 ? SQE_OK : (rc == -1 ? SQE_E_ARG : rc == -3 ? SQE_E_WORKSPACE : SQE_E_CUDA);
}

int sqe_search_gemv_sharded(const void* D, int dtype, int64_t n, int dim, const float* Q_raw, int nq, int k,
                            float* out_score, int64_t* out_idx, int64_t idx_offset, int rank, int world,
                            void* const* peer_buffers_host, int64_t capacity_entries, uint32_t epoch,
                            int flags, void* workspace, int64_t workspace_bytes, void* stream) {
    int rc = check_common("search_gemv_sharded", D, dtype, n, dim, Q_raw, nq);
    if (rc != SQE_OK) return rc;
    if (k < 1 || k > SQE_MAX_K_GEMV) { set_error("search_gemv_sharded: k=%d not in [1,%d]", k, SQE_MAX_K_GEMV); return SQE_E_ARG; }
    if (nq == 0) return SQE_OK;
    if (!out_score || !out_idx || !workspace) { set_error("search_gemv_sharded: null output/workspace"); return SQE_E_ARG; }
    XchgArgs x;
    if (make_xchg_args(&x, rank, world, peer_buffers_host, capacity_entries, epoch, nq, k) != 0) return SQE_E_ARG;
    DevInfo d;
    rc = device_info(&d);
    if (rc != SQE_OK) return rc;
    rc = launch_topk_gemv(D, dtype, n, Q_raw, true, nq, k, out_score, out_idx, idx_offset, workspace,
                          workspace_bytes, d.sm_count, static_cast<cudaStream_t>(stream), &x,
                          (flags & SQE_FLAG_QUERIES_READY) != 0);
    return rc == 0 ? SQE_OK : (rc == -1 ? SQE_E_ARG : rc == -3 ? SQE_E_WORKSPACE : SQE_E_CUDA);
}

int sqe_search_gemv_prefiltered(const void* D, int dtype, int64_t n, int dim, const void* D8, const void* meta,
                                const float* Q_raw, int nq, int k, float* out_score, int64_t* out_idx,
                                int64_t idx_offset, uint32_t* out_rescored, int rank, int world,
                                void* const* peer_buffers_host, int64_t capacity_entries, uint32_t epoch,
                                int flags, void* workspace, int64_t workspace_bytes, void* stream) {
    int rc = check_common("search_gemv_prefiltered", D, dtype, n, dim, Q_raw, nq);
    if (rc != SQE_OK) return rc;
    if (k < 1 || k > SQE_MAX_K_GEMV) { set_error("search_gemv_prefiltered: k=%d not in [1,%d]", k, SQE_MAX_K_GEMV); return SQE_E_ARG; }
    if (nq > SQE_MAX_NQ_PREFILTER) { set_error("search_gemv_prefiltered: nq=%d > %d", nq, SQE_MAX_NQ_PREFILTER); return SQE_E_ARG; }
    if (nq == 0) return SQE_OK;
    if (!out_score || !out_idx || !workspace) { set_error("search_gemv_prefiltered: null output/workspace"); return SQE_E_ARG; }
    if (n > 0 && (!D8 || !meta || !aligned16(D8) || !aligned16(meta))) {
        set_error("search_gemv_prefiltered: null or unaligned coarse rows");
        return SQE_E_ARG;
    }
    XchgArgs x;
    if (make_xchg_args(&x, rank, world, peer_buffers_host, capacity_entries, epoch, nq, k) != 0) return SQE_E_ARG;
    DevInfo d;
    rc = device_info(&d);
    if (rc != SQE_OK) return rc;
    rc = launch_topk_prefiltered(D, dtype, n, D8, meta, Q_raw, true, nq, k, out_score, out_idx, idx_offset,
                                 out_rescored, workspace, workspace_bytes, d.sm_count,
                                 static_cast<cudaStream_t>(stream), &x, (flags & SQE_FLAG_QUERIES_READY) != 0);
    return rc == 0 ? SQE_OK : (rc == -1 ? SQE_E_ARG : rc == -3 ? SQE_E_WORKSPACE : SQE_E_CUDA);
}

int64_t sqe_search_batched_prefiltered_workspace_bytes(int64_t n, int b, int k, int dtype) {
    DevInfo d;
    if (device_info(&d) != SQE_OK) d.sm_count = 160;
    if (b < 1) b = 1;
    if (k < 1) k = 1;
    return batched_i8_workspace_bytes(n, b, k, dtype, d.sm_count);
}

int sqe_search_batched_prefiltered(const void* D, int dtype, int64_t n, int dim, const void* D8, const void* meta,
                                   const float* Q_raw, int b, int k, float* out_score, int64_t* out_idx,
                                   int64_t idx_offset, uint32_t* out_rescored, void* workspace,
                                   int64_t workspace_bytes, void* stream) {
    int rc = check_common("search_batched_prefiltered", D, dtype, n, dim, Q_raw, b);
    if (rc != SQE_OK) return rc;
    if (k < 1 || k > SQE_MAX_K_BATCHED) { set_error("search_batched_prefiltered: k=%d not in [1,%d]", k, SQE_MAX_K_BATCHED); return SQE_E_ARG; }
    if (b == 0) return SQE_OK;
    if (!out_score || !out_idx || !workspace) { set_error("search_batched_prefiltered: null output/workspace"); return SQE_E_ARG; }
    if (n > 0 && (!D8 || !meta || !aligned16(D8) || !aligned16(meta))) {
        set_error("search_batched_prefiltered: null or unaligned coarse rows");
        return SQE_E_ARG;
    }
    if (!aligned16(workspace)) { set_error("search_batched_prefiltered: workspace must be 16-byte aligned"); return SQE_E_ARG; }
    DevInfo d;
    rc = device_info(&d);
    if (rc != SQE_OK) return rc;
    rc = launch_search_batched_prefiltered(D, dtype, n, D8, meta, Q_raw, b, k, out_score, out_idx, idx_offset,
                                           out_rescored, workspace, workspace_bytes, d.sm_count,
                                           static_cast<cudaStream_t>(stream));
    return rc == 0 ? SQE_OK : (rc == -1 ? SQE_E_ARG : rc == -3 ? SQE_E_WORKSPACE : SQE_E_CUDA);
}

int64_t sqe_topk_batched_workspace_bytes(int64_t n, int b, int k) {
    DevInfo d;
    if (device_info(&d) != SQE_OK) d.sm_count = 160;
    if (b < 1) b = 1;
    if (k < 1) k = 1;
    return batched_workspace_bytes(n, b, k, d.sm_count);
}

int sqe_topk_batched(const void* D, int dtype, int64_t n, int dim, const void* Q, int b, int k,
                     float* out_score, int64_t* out_idx, int64_t idx_offset, void* workspace,
                     int64_t workspace_bytes, void* stream) {
    int rc = check_common("topk_batched", D, dtype, n, dim, Q, b);
    if (rc != SQE_OK) return rc;
    if (dtype == SQE_F32) { set_error("topk_batched: fp32 shards use sqe_topk_gemv (tensor path is bf16/fp16/split bf16)"); return SQE_E_UNSUPPORTED; }
    if (k < 1 || k > SQE_MAX_K_BATCHED) { set_error("topk_batched: k=%d not in [1,%d]", k, SQE_MAX_K_BATCHED); return SQE_E_ARG; }
    if (b == 0) return SQE_OK;
    if (!out_score || !out_idx || !workspace) { set_error("topk_batched: null output/workspace"); return SQE_E_ARG; }
    DevInfo d;
    rc = device_info(&d);
    if (rc != SQE_OK) return rc;
    return launch_topk_batched(D, dtype, n, Q, b, k, out_score, out_idx, idx_offset, workspace,
                               workspace_bytes, d.sm_count, static_cast<cudaStream_t>(stream));
}

// workspace of cache_top1 = [b] fp32 score + [b] int64 idx staging, then the scorer's own
static int64_t cache_stage_bytes(int b) { return ((static_cast<int64_t>(b) * 12 + 255) / 256) * 256; }

int64_t sqe_cache_top1_workspace_bytes(int64_t n, int b) {
    if (b < 1) b = 1;
    int64_t g = sqe_topk_gemv_workspace_bytes(b, 1);
    int64_t t = sqe_topk_batched_workspace_bytes(n, b, 1);
    return cache_stage_bytes(b) + 256 + (g > t ? g : t);      // + alignment slack for the staging area
}

int sqe_cache_top1(const void* C, int dtype, int64_t n, int dim, const void* Q, int b,
                   double threshold, float* out_score, int32_t* out_idx, uint8_t* out_hit, int path,
                   void* workspace, int64_t workspace_bytes, void* stream) {
    int rc = check_common("cache_top1", C, dtype, n, dim, Q, b);
    if (rc != SQE_OK) return rc;
    if (b == 0) return SQE_OK;
    if (!out_score || !out_idx || !out_hit || !workspace) { set_error("cache_top1: null output/workspace"); return SQE_E_ARG; }
    if (path < 0 || path > 2) { set_error("cache_top1: bad path %d", path); return SQE_E_ARG; }
    const int64_t stage = cache_stage_bytes(b);
    if (workspace_bytes < stage) { set_error("cache_top1: workspace too small"); return SQE_E_WORKSPACE; }
    // Layout: [kernel workspace (zeroed 4 KB header first) ... | staging (idx, score) at the END].
    // The header must sit at the same address for every b and every path: it is shared state
    // between calls (GEMV ticket counters; both kernels leave it zero).
    char* ws = static_cast<char*>(workspace);
    char* st = ws + ((workspace_bytes - stage) & ~static_cast<int64_t>(255));
    int64_t* st_idx = reinterpret_cast<int64_t*>(st);
    float* st_score = reinterpret_cast<float*>(st + static_cast<int64_t>(b) * 8);
    const int64_t kernel_bytes = st - ws;
    const bool tensor = (path == 2) || (path == 0 && dtype != SQE_F32 && b > 1);
    if (tensor && dtype == SQE_F32) { set_error("cache_top1: tensor path needs a bf16/fp16 cache"); return SQE_E_UNSUPPORTED; }
    if (tensor)
        rc = sqe_topk_batched(C, dtype, n, dim, Q, b, 1, st_score, st_idx, 0, ws, kernel_bytes, stream);
    else
        rc = sqe_topk_gemv(C, dtype, n, dim, Q, b, 1, st_score, st_idx, 0, ws, kernel_bytes, stream);
    if (rc != SQE_OK) return rc;
    rc = launch_cache_finalize(st_score, st_idx, b, threshold, out_score, out_idx, out_hit,
                               static_cast<cudaStream_t>(stream));
    return rc == 0 ? SQE_OK : SQE_E_CUDA;
}

int64_t sqe_cache_top1_prefiltered_workspace_bytes(int64_t n, int b, int dtype) {
    if (b < 1) b = 1;
    return cache_stage_bytes(b) + 256 + sqe_search_batched_prefiltered_workspace_bytes(n, b, 1, dtype);
}

int sqe_cache_top1_prefiltered(const void* C, int dtype, int64_t n, int dim, const void* C8, const void* meta,
                               const float* Q_raw, int b, double threshold, float* out_score, int32_t* out_idx,
                               uint8_t* out_hit, void* workspace, int64_t workspace_bytes, void* stream) {
    int rc = check_common("cache_top1_prefiltered", C, dtype, n, dim, Q_raw, b);
    if (rc != SQE_OK) return rc;
    if (b == 0) return SQE_OK;
    if (!out_score || !out_idx || !out_hit || !workspace) { set_error("cache_top1_prefiltered: null output/workspace"); return SQE_E_ARG; }
    const int64_t stage = cache_stage_bytes(b);
    if (workspace_bytes < stage) { set_error("cache_top1_prefiltered: workspace too small"); return SQE_E_WORKSPACE; }
    // same layout as sqe_cache_top1: [kernel workspace ... | staging (idx, score) at the END]
    char* ws = static_cast<char*>(workspace);
    char* st = ws + ((workspace_bytes - stage) & ~static_cast<int64_t>(255));
    int64_t* st_idx = reinterpret_cast<int64_t*>(st);
    float* st_score = reinterpret_cast<float*>(st + static_cast<int64_t>(b) * 8);
    rc = sqe_search_batched_prefiltered(C, dtype, n, dim, C8, meta, Q_raw, b, 1, st_score, st_idx, 0, nullptr, ws,
                                        st - ws, stream);
    if (rc != SQE_OK) return rc;
    rc = launch_cache_finalize(st_score, st_idx, b, threshold, out_score, out_idx, out_hit,
                               static_cast<cudaStream_t>(stream));
    return rc == 0 ? SQE_OK : SQE_E_CUDA;
}

int sqe_merge_topk(const float* scores, const int64_t* idx, int lists, int b, int k_in, int k_out,
                   float* out_score, int64_t* out_idx, void* stream) {
    if (lists < 1 || b < 0 || k_in < 1 || k_out < 1 || k_in > SQE_MAX_K_GEMV || k_out > SQE_MAX_K_GEMV) {
        set_error("merge_topk: bad sizes lists=%d b=%d k_in=%d k_out=%d", lists, b, k_in, k_out);
        return SQE_E_ARG;
    }
    if (b == 0) return SQE_OK;
    if (!scores || !idx || !out_score || !out_idx) { set_error("merge_topk: null pointer"); return SQE_E_ARG; }
    DevInfo d;
    int rc = device_info(&d);
    if (rc != SQE_OK) return rc;
    rc = launch_merge_topk(scores, idx, lists, b, k_in, k_out, out_score, out_idx,
                           static_cast<cudaStream_t>(stream));
    return rc == 0 ? SQE_OK : SQE_E_CUDA;
}

int64_t sqe_exchange_buffer_bytes(int world, int64_t capacity_entries) {
    if (world < 1 || capacity_entries < 0) return SQE_E_ARG;
    return exchange_buffer_bytes(world, capacity_entries);
}

int sqe_exchange_merge(const float* scores, const int64_t* idx, int b, int k_in, int k_out, int rank,
                       int world, void* const* peer_buffers_host, int64_t capacity_entries,
                       uint32_t epoch, uint32_t wait_mask, float* out_score, int64_t* out_idx,
                       void* stream) {
    if (world < 1 || world > 16 || rank < 0 || rank >= world || b < 0 || k_in < 1 || k_out < 1 ||
        k_in > SQE_MAX_K_GEMV || k_out > SQE_MAX_K_GEMV ||
        capacity_entries < static_cast<int64_t>(b) * k_in) {
        set_error("exchange_merge: bad sizes world=%d rank=%d b=%d k_in=%d k_out=%d cap=%lld", world,
                  rank, b, k_in, k_out, (long long)capacity_entries);
        return SQE_E_ARG;
    }
    if (b == 0) return SQE_OK;
    if (!scores || !idx || !out_score || !out_idx || !peer_buffers_host) { set_error("exchange_merge: null pointer"); return SQE_E_ARG; }
    for (int g = 0; g < world; ++g)
        if (!peer_buffers_host[g] || !aligned16(peer_buffers_host[g])) { set_error("exchange_merge: peer buffer %d null or unaligned", g); return SQE_E_ARG; }
    DevInfo d;
    int rc = device_info(&d);
    if (rc != SQE_OK) return rc;
    rc = launch_exchange_merge(scores, idx, b, k_in, k_out, rank, world, peer_buffers_host,
                               capacity_entries, epoch, wait_mask, out_score, out_idx, d.sm_count,
                               static_cast<cudaStream_t>(stream));
    return rc == 0 ? SQE_OK : (rc == -1 ? SQE_E_ARG : SQE_E_CUDA);
}

// ------------------------------------------------------------------ ENC  embedding encoder
static int rc_map(int rc) { return rc == 0 ? SQE_OK : (rc == -1 ? SQE_E_ARG : SQE_E_CUDA); }

int sqe_encoder_embed_ln(const int32_t* ids, const int32_t* pos, const float* word_emb, int vocab,
                         const float* pos_emb, int max_pos, const float* type_emb, const float* gamma,
                         const float* beta, float eps, int64_t rows, float* out_f32, void* out_f16, float* stats,
                         void* stream) {
    if (rows < 0 || vocab < 1 || max_pos < 1) { set_error("encoder_embed_ln: bad sizes"); return SQE_E_ARG; }
    if (rows == 0) return SQE_OK;
    if (!ids || !pos || !word_emb || !pos_emb || !type_emb || !gamma || !beta || !out_f32 || !out_f16 ||
        !aligned16(word_emb) || !aligned16(pos_emb) || !aligned16(type_emb) || !aligned16(gamma) || !aligned16(beta) ||
        !aligned16(out_f32) || !aligned16(out_f16)) {
        set_error("encoder_embed_ln: null or unaligned pointer");
        return SQE_E_ARG;
    }
    DevInfo d;
    int rc = device_info(&d);
    if (rc != SQE_OK) return rc;
    return rc_map(launch_encoder_embed_ln(ids, pos, word_emb, pos_emb, type_emb, gamma, beta, eps, rows, vocab, max_pos,
                                          out_f32, out_f16, stats, static_cast<cudaStream_t>(stream)));
}

int sqe_encoder_layernorm(const float* in, const float* gamma, const float* beta, float eps, int64_t rows,
                          float* out_f32, void* out_f16, float* stats, void* stream) {
    if (rows < 0) { set_error("encoder_layernorm: negative rows"); return SQE_E_ARG; }
    if (rows == 0) return SQE_OK;
    if (!in || !gamma || !beta || (!out_f32 && !stats) || !out_f16 || !aligned16(in) || !aligned16(gamma) || !aligned16(beta) ||
        !aligned16(out_f32) || !aligned16(out_f16) || (reinterpret_cast<uintptr_t>(stats) & 7u) != 0) {
        set_error("encoder_layernorm: null or unaligned pointer");
        return SQE_E_ARG;
    }
    DevInfo d;
    int rc = device_info(&d);
    if (rc != SQE_OK) return rc;
    return rc_map(launch_encoder_layernorm(in, gamma, beta, eps, rows, out_f32, out_f16, stats, static_cast<cudaStream_t>(stream)));
}

int sqe_encoder_gemm(const void* X, int64_t ldx, const void* W, const float* bias, int64_t m, int n, int k,
                     int epilogue, void* out0, int64_t ld0, void* out1, int64_t ld1, int n_split, int q_cols,
                     float q_scale, const float* residual, int64_t ldr, const float* res_stats, const float* res_gamma,
                     const float* res_beta, void* stream) {
    if (m < 0 || m >= (1LL << 31) - 256 || n < 256 || n % 256 != 0 || k < 64 || k % 64 != 0) {
        set_error("encoder_gemm: need 0 <= m < 2^31, n %% 256 == 0, k %% 64 == 0 (m=%lld n=%d k=%d)", (long long)m, n, k);
        return SQE_E_ARG;
    }
    if (epilogue < SQE_ENC_EPI_SPLIT || epilogue > SQE_ENC_EPI_GELU) { set_error("encoder_gemm: bad epilogue %d", epilogue); return SQE_E_ARG; }
    if (m == 0) return SQE_OK;
    if (!X || !W || !bias || !out0 || !aligned16(X) || !aligned16(W) || !aligned16(bias) || !aligned16(out0) ||
        ldx < k || ldx % 8 != 0 || ld0 % 8 != 0) {
        set_error("encoder_gemm: null / unaligned pointer or bad leading dimension");
        return SQE_E_ARG;
    }
    if (epilogue == SQE_ENC_EPI_SPLIT) {
        if (n_split < 0 || n_split > n || n_split % 64 != 0 || q_cols < 0 || q_cols > n_split || q_cols % 32 != 0 ||
            ld0 < n_split || (n_split < n && (!out1 || ld1 < m))) {
            set_error("encoder_gemm: bad split (n_split=%d q_cols=%d ld0=%lld ld1=%lld)", n_split, q_cols, (long long)ld0,
                      (long long)ld1);
            return SQE_E_ARG;
        }
    } else if (ld0 < n) {
        set_error("encoder_gemm: ld0 < n");
        return SQE_E_ARG;
    }
    if (epilogue == SQE_ENC_EPI_RES_F32 && (!residual || !aligned16(residual) || ldr < n || ldr % 4 != 0)) {
        set_error("encoder_gemm: residual null / unaligned / too narrow");
        return SQE_E_ARG;
    }
    if (res_stats && (epilogue != SQE_ENC_EPI_RES_F32 || !res_gamma || !res_beta || !aligned16(res_gamma) || !aligned16(res_beta) ||
                      (reinterpret_cast<uintptr_t>(res_stats) & 7u) != 0)) {
        set_error("encoder_gemm: LayerNorm statistics need the residual epilogue and aligned gamma / beta");
        return SQE_E_ARG;
    }
    DevInfo d;
    int rc = device_info(&d);
    if (rc != SQE_OK) return rc;
    return rc_map(launch_encoder_gemm(X, ldx, W, bias, m, n, k, epilogue, out0, ld0, out1, ld1, n_split, q_cols, q_scale,
                                      residual, ldr, res_stats, res_gamma, res_beta, d.sm_count,
                                      static_cast<cudaStream_t>(stream)));
}

int64_t sqe_encoder_gemm_small_workspace_bytes(void) { return encoder_gemm_small_workspace_bytes(); }

int sqe_encoder_gemm_small(const void* X, int64_t ldx, const void* W, const float* bias, int64_t m, int n, int k,
                           int epilogue, void* out0, int64_t ld0, void* out1, int64_t ld1, int n_split, int q_cols,
                           float q_scale, const float* residual, int64_t ldr, const float* res_stats,
                           const float* res_gamma, const float* res_beta, void* workspace, int64_t workspace_bytes,
                           void* stream) {
    if (m == 0) return SQE_OK;
    if (res_stats && (epilogue != SQE_ENC_EPI_RES_F32 || !res_gamma || !res_beta)) {
        set_error("encoder_gemm_small: LayerNorm statistics need the residual epilogue, gamma and beta");
        return SQE_E_ARG;
    }
    if (epilogue < SQE_ENC_EPI_SPLIT || epilogue > SQE_ENC_EPI_GELU) { set_error("encoder_gemm_small: bad epilogue %d", epilogue); return SQE_E_ARG; }
    if (!X || !W || !bias || !out0 || !workspace || !aligned16(X) || !aligned16(W) || !aligned16(workspace) || ldx < k ||
        ldx % 8 != 0 || (epilogue == SQE_ENC_EPI_RES_F32 && !residual) ||
        (epilogue == SQE_ENC_EPI_SPLIT && (n_split < 0 || n_split > n || (n_split < n && (!out1 || ld1 < m))))) {
        set_error("encoder_gemm_small: null / unaligned pointer or bad leading dimension");
        return SQE_E_ARG;
    }
    DevInfo d;
    int rc = device_info(&d);
    if (rc != SQE_OK) return rc;
    rc = launch_encoder_gemm_small(X, ldx, W, bias, m, n, k, epilogue, out0, ld0, out1, ld1, n_split, q_cols, q_scale,
                                   residual, ldr, res_stats, res_gamma, res_beta, workspace, workspace_bytes,
                                   static_cast<cudaStream_t>(stream));
    if (rc == -1) { set_error("encoder_gemm_small: shape not taken (m=%lld n=%d k=%d; m <= 128, n %% 128 == 0, n <= 4096)", (long long)m, n, k); return SQE_E_UNSUPPORTED; }
    return rc == 0 ? SQE_OK : SQE_E_CUDA;
}

int sqe_encoder_attention(const void* qk, const void* vt, int64_t t_pad, const int32_t* tiles, int n_tiles,
                          int max_len, void* ctx, void* stream) {
    if (n_tiles < 0 || n_tiles > 0x7fffffff / 16 || t_pad < 0 || t_pad % 8 != 0 || t_pad >= (1LL << 31) - 1024) {
        set_error("encoder_attention: bad sizes (n_tiles=%d t_pad=%lld; t_pad %% 8 == 0)", n_tiles, (long long)t_pad);
        return SQE_E_ARG;
    }
    if (n_tiles == 0) return SQE_OK;
    if (max_len < 1 || max_len > SQE_ENC_MAX_TOKENS) { set_error("encoder_attention: max_len=%d not in [1,%d]", max_len, SQE_ENC_MAX_TOKENS); return SQE_E_ARG; }
    if (!qk || !vt || !tiles || !ctx || !aligned16(qk) || !aligned16(vt) || !aligned16(tiles) || !aligned16(ctx)) {
        set_error("encoder_attention: null or unaligned pointer");
        return SQE_E_ARG;
    }
    DevInfo d;
    int rc = device_info(&d);
    if (rc != SQE_OK) return rc;
    return rc_map(launch_encoder_attention(qk, vt, t_pad, tiles, n_tiles, max_len, ctx, static_cast<cudaStream_t>(stream)));
}

int sqe_encoder_pool(const float* h, const int32_t* first_token, int n_seq, float* out, int64_t ldo, const float* stats,
                     const float* gamma, const float* beta, void* stream) {
    if (n_seq < 0 || ldo < SQE_ENC_HIDDEN || ldo % 4 != 0) { set_error("encoder_pool: bad sizes"); return SQE_E_ARG; }
    if (n_seq == 0) return SQE_OK;
    if (!h || !first_token || !out || !aligned16(h) || !aligned16(out) || (stats && (!gamma || !beta || !aligned16(gamma) || !aligned16(beta)))) {
        set_error("encoder_pool: null or unaligned pointer");
        return SQE_E_ARG;
    }
    DevInfo d;
    int rc = device_info(&d);
    if (rc != SQE_OK) return rc;
    return rc_map(launch_encoder_pool(h, first_token, n_seq, out, ldo, stats, gamma, beta, static_cast<cudaStream_t>(stream)));
}

int sqe_encoder_forward(const SqeEncoderWeights* w, const SqeEncoderBuffers* b, const int32_t* ids, const int32_t* pos,
                        const int32_t* tiles, int n_tiles, int max_len, const int32_t* first_token, int n_seq,
                        int64_t rows_used, float* out, int64_t ldo, void* stream) {
    if (!w || !b || !w->layers || w->n_layers < 1 || w->intermediate < 256 || w->intermediate % 256 != 0) {
        set_error("encoder_forward: bad weights (layers / intermediate size)");
        return SQE_E_ARG;
    }
    const int64_t m = b->t_pad;
    if (m <= 0 || m % 128 != 0) { set_error("encoder_forward: t_pad must be a positive multiple of 128"); return SQE_E_ARG; }
    const int H = SQE_ENC_HIDDEN, I = w->intermediate;
    if (rows_used < 0 || rows_used > m) { set_error("encoder_forward: rows_used out of range"); return SQE_E_ARG; }
    // a handful of tokens (one or two queries): the swap-AB split-K form of the four products; its last-CTA
    // reduction grows with the token count, beyond 32 rows the 128 x 64 tiles are as fast
    const bool small = g_enc_small.load() == 0 && rows_used > 0 && rows_used <= 32 && b->small_ws != nullptr &&
                       b->small_ws_bytes >= encoder_gemm_small_workspace_bytes() && I <= 4096;
    const int64_t mg = small ? (rows_used + 15) / 16 * 16 : m;        // whole 16-row groups (filler rows are zeros)
    // LayerNorm statistics form: a LayerNorm stores its fp16 output (the next operand) and {mean, rstd} per row;
    // its fp32 output, needed only as the next residual, is recomputed in that GEMM's epilogue from the
    // pre-LayerNorm sum (bit-identical).  sum_b = the sum the current residual comes from, sum_a the other buffer.
    auto gemm = [&](const void* X, int64_t ldx, const void* W, const float* bias, int n, int k, int epi, void* out0,
                    int64_t ld0, void* out1, int64_t ld1, int n_split, int q_cols, float q_scale, const float* res,
                    const float* rstats, const float* rgamma, const float* rbeta) {
        if (small)
            return sqe_encoder_gemm_small(X, ldx, W, bias, mg, n, k, epi, out0, ld0, out1, ld1, n_split, q_cols, q_scale,
                                          res, H, rstats, rgamma, rbeta, b->small_ws, b->small_ws_bytes, stream);
        return sqe_encoder_gemm(X, ldx, W, bias, mg, n, k, epi, out0, ld0, out1, ld1, n_split, q_cols, q_scale, res, H,
                                rstats, rgamma, rbeta, stream);
    };
    if (!b->sum_a || !b->sum_b || !b->stats_a || !b->stats_b) { set_error("encoder_forward: null buffer"); return SQE_E_ARG; }
    int rc = sqe_encoder_embed_ln(ids, pos, w->word_emb, w->vocab, w->pos_emb, w->max_pos, w->type_emb, w->emb_gamma,
                                  w->emb_beta, w->eps, m, b->sum_b, b->h16, b->stats_b, stream);
    const float* pg = w->emb_gamma;                            // the LayerNorm whose output is the current residual
    const float* pb = w->emb_beta;
    for (int l = 0; l < w->n_layers && rc == SQE_OK; ++l) {
        const SqeEncoderLayer& L = w->layers[l];
        rc = gemm(b->h16, H, L.wqkv, L.bqkv, 3 * H, H, SQE_ENC_EPI_SPLIT, b->qk, 2 * H, b->vt, m, 2 * H, H, 0.125f, nullptr,
                  nullptr, nullptr, nullptr);
        if (rc == SQE_OK) rc = sqe_encoder_attention(b->qk, b->vt, m, tiles, n_tiles, max_len, b->ctx, stream);
        if (rc == SQE_OK)
            rc = gemm(b->ctx, H, L.wo, L.bo, H, H, SQE_ENC_EPI_RES_F32, b->sum_a, H, nullptr, 0, 0, 0, 1.0f, b->sum_b,
                      b->stats_b, pg, pb);
        if (rc == SQE_OK)
            rc = sqe_encoder_layernorm(b->sum_a, L.ln1_gamma, L.ln1_beta, w->eps, m, nullptr, b->h16, b->stats_a, stream);
        if (rc == SQE_OK)
            rc = gemm(b->h16, H, L.w1, L.b1, I, H, SQE_ENC_EPI_GELU, b->ffn, I, nullptr, 0, 0, 0, 1.0f, nullptr, nullptr,
                      nullptr, nullptr);
        if (rc == SQE_OK)
            rc = gemm(b->ffn, I, L.w2, L.b2, H, I, SQE_ENC_EPI_RES_F32, b->sum_b, H, nullptr, 0, 0, 0, 1.0f, b->sum_a,
                      b->stats_a, L.ln1_gamma, L.ln1_beta);
        if (rc == SQE_OK)
            rc = sqe_encoder_layernorm(b->sum_b, L.ln2_gamma, L.ln2_beta, w->eps, m, nullptr, b->h16, b->stats_b, stream);
        pg = L.ln2_gamma;
        pb = L.ln2_beta;
    }
    if (rc == SQE_OK) rc = sqe_encoder_pool(b->sum_b, first_token, n_seq, out, ldo, b->stats_b, pg, pb, stream);
    return rc;
}

}  // extern "C"
