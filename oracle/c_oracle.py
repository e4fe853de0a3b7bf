"""ctypes wrapper of oracle/c_oracle.c -- the plain-C restatement of the reference path.
TEST INFRASTRUCTURE ONLY (see oracle/__init__.py).  `build()` compiles it with gcc; the shared
object goes to oracle/_build/ (git-ignored, travels with gpurun snapshots)."""
from __future__ import annotations

import ctypes
import os
import subprocess
from typing import Tuple

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
SRC = os.path.join(HERE, "c_oracle.c")
LIB = os.path.join(HERE, "_build", "libsqe_oracle.so")
DIM = 1024
_lib = None


def build(force: bool = False) -> str:
    """gcc -O2 -ffp-contract=off (no FMA contraction: every fp32 operation rounds once)."""
    if not force and os.path.isfile(LIB) and os.path.getmtime(LIB) >= os.path.getmtime(SRC):
        return LIB
    os.makedirs(os.path.dirname(LIB), exist_ok=True)
    subprocess.run(["gcc", "-O2", "-ffp-contract=off", "-fno-fast-math", "-fPIC", "-shared", "-o", LIB, SRC, "-lm"],
                   check=True)
    return LIB


def available() -> bool:
    return os.path.isfile(LIB) or os.path.isfile(SRC) and _have_gcc()


def _have_gcc() -> bool:
    from shutil import which
    return which("gcc") is not None


def _load():
    global _lib
    if _lib is None:
        path = build() if _have_gcc() else LIB
        lib = ctypes.CDLL(path)
        f32p, i64p = ctypes.POINTER(ctypes.c_float), ctypes.POINTER(ctypes.c_int64)
        lib.sqe_oracle_row_sumsq.restype = ctypes.c_float
        lib.sqe_oracle_row_sumsq.argtypes = [f32p]
        lib.sqe_oracle_normalize_rows.argtypes = [f32p, f32p, ctypes.c_int64]
        lib.sqe_oracle_cosine.restype = ctypes.c_double
        lib.sqe_oracle_cosine.argtypes = [f32p, f32p]
        lib.sqe_oracle_scores.argtypes = [f32p, ctypes.c_int64, f32p, ctypes.c_int, f32p]
        lib.sqe_oracle_topk.argtypes = [f32p, ctypes.c_int, ctypes.c_int64, ctypes.c_int, f32p, i64p]
        lib.sqe_oracle_cache_lookup.argtypes = [f32p, ctypes.c_int64, f32p, ctypes.c_double,
                                                ctypes.POINTER(ctypes.c_int32), f32p,
                                                ctypes.POINTER(ctypes.c_uint8)]
        lib.sqe_oracle_quantize_row.argtypes = [f32p, ctypes.POINTER(ctypes.c_int8), f32p]
        lib.sqe_oracle_prefilter_bounds.argtypes = [f32p, ctypes.c_int64, f32p, ctypes.c_int, f32p, f32p]
        _lib = lib
    return _lib


def _f32(a: np.ndarray):
    return a.ctypes.data_as(ctypes.POINTER(ctypes.c_float))


def _rows(x) -> np.ndarray:
    a = np.ascontiguousarray(x, dtype=np.float32)
    if a.ndim == 1:
        a = a[None, :]
    assert a.ndim == 2 and a.shape[1] == DIM, a.shape
    return a


def row_sumsq(emb) -> np.ndarray:
    a = _rows(emb)
    lib = _load()
    return np.array([lib.sqe_oracle_row_sumsq(_f32(a[i])) for i in range(a.shape[0])], dtype=np.float32)


def normalize_rows(emb) -> np.ndarray:
    a = _rows(emb)
    out = np.empty_like(a)
    _load().sqe_oracle_normalize_rows(_f32(a), _f32(out), a.shape[0])
    return out


def cosine_similarity(a, b) -> float:
    x, y = _rows(a), _rows(b)
    return float(_load().sqe_oracle_cosine(_f32(x), _f32(y)))


def topk_cosine(d_stored, q_stored, k: int) -> Tuple[np.ndarray, np.ndarray]:
    """Exact scoring of stored rows + best-first top-k, ties -> lower row."""
    d, q = _rows(d_stored), _rows(q_stored)
    n, b = d.shape[0], q.shape[0]
    scores = np.empty((b, n), dtype=np.float32)
    lib = _load()
    lib.sqe_oracle_scores(_f32(d), n, _f32(q), b, _f32(scores))
    return topk_from_scores(scores, k)


def topk_from_scores(scores, k: int) -> Tuple[np.ndarray, np.ndarray]:
    s = np.ascontiguousarray(scores, dtype=np.float32)
    b, n = s.shape
    os_ = np.empty((b, k), dtype=np.float32)
    oi = np.empty((b, k), dtype=np.int64)
    _load().sqe_oracle_topk(_f32(s), b, n, k, _f32(os_), oi.ctypes.data_as(ctypes.POINTER(ctypes.c_int64)))
    return os_, oi


def cache_lookup(c_stored, q_stored_row, threshold: float) -> Tuple[int, float, bool]:
    c, q = _rows(c_stored), _rows(q_stored_row)
    idx, sim, hit = ctypes.c_int32(), ctypes.c_float(), ctypes.c_uint8()
    _load().sqe_oracle_cache_lookup(_f32(c), c.shape[0], _f32(q), float(threshold),
                                    ctypes.byref(idx), ctypes.byref(sim), ctypes.byref(hit))
    return int(idx.value), float(sim.value), bool(hit.value)


def quantize_rows_int8(x) -> Tuple[np.ndarray, np.ndarray]:
    """The K3p quantiser in plain C: (int8 [n,1024], fp32 [n,4] = {sd, eps, nd, 0})."""
    a = _rows(x)
    d8 = np.empty(a.shape, dtype=np.int8)
    meta = np.empty((a.shape[0], 4), dtype=np.float32)
    lib = _load()
    for i in range(a.shape[0]):
        lib.sqe_oracle_quantize_row(_f32(a[i]), d8[i].ctypes.data_as(ctypes.POINTER(ctypes.c_int8)), _f32(meta[i]))
    return d8, meta


def prefilter_bounds(d_stored, q_stored) -> Tuple[np.ndarray, np.ndarray]:
    """(L, U) [B,N]: lower / upper bounds of every exact score, as the int8 prefilter derives them."""
    d, q = _rows(d_stored), _rows(q_stored)
    n, b = d.shape[0], q.shape[0]
    L = np.empty((b, n), dtype=np.float32)
    U = np.empty((b, n), dtype=np.float32)
    _load().sqe_oracle_prefilter_bounds(_f32(d), n, _f32(q), b, _f32(L), _f32(U))
    return L, U
