"""ctypes binding of include/sqe_b200.h.  The only way into the CUDA kernels.

There is deliberately no fallback: if the shared library has not been built
(`python semantic-query-engine_b200/build.py`, or `__graft_entry__.build()`),
importing is fine but the first call raises `NativeLibraryMissing`.
"""
from __future__ import annotations

import ctypes
import os
import threading
from ctypes import POINTER, c_char_p, c_double, c_float, c_int, c_int32, c_int64, c_uint8, c_uint32, c_void_p

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(HERE, "lib", "libsqe_b200.so")

SQE_F32, SQE_BF16, SQE_F16, SQE_BF16X2 = 0, 1, 2, 3
SQE_DIM = 1024
SQE_MAX_K_GEMV = 256
SQE_MAX_K_BATCHED = 128
SQE_MAX_NQ_PREFILTER = 64
SQE_MAX_NQ_FUSED_EXCHANGE = 2
SQE_FLAG_QUERIES_READY = 1
DTYPE_CODES = {"fp32": SQE_F32, "bf16": SQE_BF16, "fp16": SQE_F16, "bf16x2": SQE_BF16X2}

# every symbol include/sqe_b200.h declares: (name, restype, argtypes)
PROTOTYPES = [
    ("sqe_abi_version", c_int, []),
    ("sqe_last_error", c_char_p, []),
    ("sqe_device_info", c_int, [POINTER(c_int), POINTER(c_int), POINTER(c_int)]),
    ("sqe_tuning_set", c_int, [c_int, c_int]),
    ("sqe_debug_k2_timers", None, [c_void_p]),
    ("sqe_debug_encoder_attention_timers", None, [c_void_p]),
    ("sqe_debug_encoder_gemm_timers", None, [c_void_p]),
    ("sqe_normalize_cast", c_int, [c_void_p, c_void_p, c_int64, c_int, c_int, c_void_p]),
    ("sqe_topk_gemv_workspace_bytes", c_int64, [c_int, c_int]),
    ("sqe_topk_gemv", c_int, [c_void_p, c_int, c_int64, c_int, c_void_p, c_int, c_int,
                              c_void_p, c_void_p, c_int64, c_void_p, c_int64, c_void_p]),
    ("sqe_search_gemv", c_int, [c_void_p, c_int, c_int64, c_int, c_void_p, c_int, c_int,
                                c_void_p, c_void_p, c_int64, c_void_p, c_int64, c_void_p]),
    ("sqe_quantize_rows", c_int, [c_void_p, c_int, c_int64, c_int, c_void_p, c_void_p, c_void_p]),
    ("sqe_topk_gemv_prefiltered_workspace_bytes", c_int64, [c_int64, c_int, c_int]),
    ("sqe_topk_gemv_prefiltered", c_int, [c_void_p, c_int, c_int64, c_int, c_void_p, c_void_p, c_void_p,
                                          c_int, c_int, c_void_p, c_void_p, c_int64, c_void_p,
                                          c_void_p, c_int64, c_void_p]),
    ("sqe_topk_batched_workspace_bytes", c_int64, [c_int64, c_int, c_int]),
    ("sqe_topk_batched", c_int, [c_void_p, c_int, c_int64, c_int, c_void_p, c_int, c_int,
                                 c_void_p, c_void_p, c_int64, c_void_p, c_int64, c_void_p]),
    ("sqe_cache_top1_workspace_bytes", c_int64, [c_int64, c_int]),
    ("sqe_cache_top1", c_int, [c_void_p, c_int, c_int64, c_int, c_void_p, c_int, c_double,
                               c_void_p, c_void_p, c_void_p, c_int, c_void_p, c_int64, c_void_p]),
    ("sqe_merge_topk", c_int, [c_void_p, c_void_p, c_int, c_int, c_int, c_int,
                               c_void_p, c_void_p, c_void_p]),
    ("sqe_exchange_buffer_bytes", c_int64, [c_int, c_int64]),
    ("sqe_exchange_merge", c_int, [c_void_p, c_void_p, c_int, c_int, c_int, c_int, c_int,
                                   POINTER(c_void_p), c_int64, c_uint32, c_uint32,
                                   c_void_p, c_void_p, c_void_p]),
    ("sqe_search_batched_prefiltered_workspace_bytes", c_int64, [c_int64, c_int, c_int, c_int]),
    ("sqe_search_batched_prefiltered", c_int, [c_void_p, c_int, c_int64, c_int, c_void_p, c_void_p, c_void_p,
                                               c_int, c_int, c_void_p, c_void_p, c_int64, c_void_p,
                                               c_void_p, c_int64, c_void_p]),
    ("sqe_cache_top1_prefiltered_workspace_bytes", c_int64, [c_int64, c_int, c_int]),
    ("sqe_cache_top1_prefiltered", c_int, [c_void_p, c_int, c_int64, c_int, c_void_p, c_void_p, c_void_p, c_int,
                                           c_double, c_void_p, c_void_p, c_void_p, c_void_p, c_int64, c_void_p]),
    ("sqe_search_gemv_sharded", c_int, [c_void_p, c_int, c_int64, c_int, c_void_p, c_int, c_int,
                                        c_void_p, c_void_p, c_int64, c_int, c_int, POINTER(c_void_p),
                                        c_int64, c_uint32, c_int, c_void_p, c_int64, c_void_p]),
    ("sqe_search_gemv_prefiltered", c_int, [c_void_p, c_int, c_int64, c_int, c_void_p, c_void_p, c_void_p,
                                            c_int, c_int, c_void_p, c_void_p, c_int64, c_void_p,
                                            c_int, c_int, POINTER(c_void_p), c_int64, c_uint32, c_int,
                                            c_void_p, c_int64, c_void_p]),
    ("sqe_encoder_embed_ln", c_int, [c_void_p, c_void_p, c_void_p, c_int, c_void_p, c_int, c_void_p, c_void_p,
                                     c_void_p, c_float, c_int64, c_void_p, c_void_p, c_void_p, c_void_p]),
    ("sqe_encoder_layernorm", c_int, [c_void_p, c_void_p, c_void_p, c_float, c_int64, c_void_p, c_void_p, c_void_p,
                                      c_void_p]),
    ("sqe_encoder_gemm", c_int, [c_void_p, c_int64, c_void_p, c_void_p, c_int64, c_int, c_int, c_int, c_void_p,
                                 c_int64, c_void_p, c_int64, c_int, c_int, c_float, c_void_p, c_int64, c_void_p,
                                 c_void_p, c_void_p, c_void_p]),
    ("sqe_encoder_attention", c_int, [c_void_p, c_void_p, c_int64, c_void_p, c_int, c_int, c_void_p, c_void_p]),
    ("sqe_encoder_pool", c_int, [c_void_p, c_void_p, c_int, c_void_p, c_int64, c_void_p, c_void_p, c_void_p,
                                 c_void_p]),
    ("sqe_encoder_forward", c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_int, c_int, c_void_p, c_int,
                                    c_int64, c_void_p, c_int64, c_void_p]),
    ("sqe_encoder_gemm_small_workspace_bytes", c_int64, []),
    ("sqe_encoder_gemm_small", c_int, [c_void_p, c_int64, c_void_p, c_void_p, c_int64, c_int, c_int, c_int, c_void_p,
                                       c_int64, c_void_p, c_int64, c_int, c_int, c_float, c_void_p, c_int64, c_void_p,
                                       c_void_p, c_void_p, c_void_p, c_int64, c_void_p]),
]


class SqeEncoderLayer(ctypes.Structure):
    _fields_ = [(n, c_void_p) for n in ("wqkv", "bqkv", "wo", "bo", "ln1_gamma", "ln1_beta", "w1", "b1", "w2", "b2",
                                        "ln2_gamma", "ln2_beta")]


class SqeEncoderWeights(ctypes.Structure):
    _fields_ = [("n_layers", c_int), ("vocab", c_int), ("max_pos", c_int), ("intermediate", c_int), ("eps", c_float),
                ("word_emb", c_void_p), ("pos_emb", c_void_p), ("type_emb", c_void_p), ("emb_gamma", c_void_p),
                ("emb_beta", c_void_p), ("layers", POINTER(SqeEncoderLayer))]


class SqeEncoderBuffers(ctypes.Structure):
    _fields_ = [("t_pad", c_int64)] + [(n, c_void_p) for n in ("sum_a", "sum_b", "stats_a", "stats_b", "h16", "qk", "vt",
                                                                 "ctx", "ffn", "small_ws")] + [("small_ws_bytes", c_int64)]



class NativeLibraryMissing(RuntimeError):
    pass


class SqeError(RuntimeError):
    def __init__(self, code: int, message: str):
        super().__init__(f"sqe_b200 error {code}: {message}")
        self.code = code
        self.message = message


_lib = None
_lock = threading.Lock()
# kernel launches issued through this binding (bench.py reports them as gpu_launches)
launch_count = 0
LAUNCHES_PER_CALL = {
    "sqe_normalize_cast": 1,
    "sqe_topk_gemv": 1,
    "sqe_search_gemv": 1,
    "sqe_quantize_rows": 1,
    "sqe_topk_gemv_prefiltered": 2,
    "sqe_topk_batched": 2,
    "sqe_cache_top1": 2,      # +1 when it takes the tensor path (counted by the caller)
    "sqe_merge_topk": 1,
    "sqe_exchange_merge": 1,
    "sqe_search_gemv_sharded": 1,
    "sqe_search_gemv_prefiltered": 2,
    "sqe_search_batched_prefiltered": 3,      # prepare queries, int8 tensor-core scan, exact rescoring
    "sqe_cache_top1_prefiltered": 4,          # the same + the threshold epilogue
    "sqe_encoder_embed_ln": 1,
    "sqe_encoder_layernorm": 1,
    "sqe_encoder_gemm": 1,
    "sqe_encoder_attention": 1,
    "sqe_encoder_pool": 1,
    "sqe_encoder_gemm_small": 1,
    "sqe_encoder_forward": 0,                 # 2 + 7 n_layers, counted by the caller
}
SQE_ENC_EPI_SPLIT, SQE_ENC_EPI_RES_F32, SQE_ENC_EPI_GELU = 0, 1, 2
SQE_ENC_MAX_TOKENS = 512


def load():
    """Load the library (once) and attach prototypes."""
    global _lib
    if _lib is not None:
        return _lib
    with _lock:
        if _lib is not None:
            return _lib
        if not os.path.isfile(LIB_PATH):
            raise NativeLibraryMissing(
                f"{LIB_PATH} not found: build it with `python {os.path.join(HERE, 'build.py')}` "
                "(there is no CPU fallback)")
        lib = ctypes.CDLL(LIB_PATH)
        for name, restype, argtypes in PROTOTYPES:
            fn = getattr(lib, name)          # AttributeError if the ABI is incomplete
            fn.restype = restype
            fn.argtypes = argtypes
        _lib = lib
    return _lib


def last_error() -> str:
    msg = load().sqe_last_error()
    return msg.decode("utf-8", "replace") if msg else ""


def call(name: str, *args) -> None:
    """Call an int-returning entry point; raise SqeError on a non-zero status."""
    global launch_count
    rc = getattr(load(), name)(*args)
    if rc != 0:
        raise SqeError(rc, last_error())
    launch_count += LAUNCHES_PER_CALL.get(name, 0)


SQE_TUNE_K2_CTA_GROUP = 0
SQE_TUNE_K2_EPILOGUE_MODE = 1      # diagnostics only
SQE_TUNE_K2_D_HINT = 2
SQE_TUNE_K2_WINDOW = 3
SQE_TUNE_ENC_SMALL = 5         # 0 = few-token passes take the swap-AB split-K GEMM (default), 1 = never
SQE_TUNE_ENC_GEMM_FORM = 4     # 0 auto, 1 = 128 x 64 tiles, 2 = 256 x 256 tiles on CTA pairs


def tuning_set(knob: int, value: int) -> int:
    """Set a tuning knob, return its previous value."""
    old = load().sqe_tuning_set(knob, value)
    if old < 0:
        raise SqeError(old, last_error())
    return old


def device_info():
    sm, maj, mnr = c_int(0), c_int(0), c_int(0)
    rc = load().sqe_device_info(ctypes.byref(sm), ctypes.byref(maj), ctypes.byref(mnr))
    return rc, sm.value, maj.value, mnr.value
