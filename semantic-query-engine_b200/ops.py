"""Thin tensor-level wrappers over the C ABI.

torch is used here only to own device memory and streams; every computation is
a call into libsqe_b200.so (hand-written sm_100a kernels).  All functions take
CUDA tensors, launch on the current stream of the tensor's device and return
CUDA tensors without synchronising.
"""
from __future__ import annotations

import ctypes
import queue
import threading as _threading
from typing import Optional, Tuple

import numpy as np
import torch

from . import _native as nat

TORCH_DTYPES = {"fp32": torch.float32, "bf16": torch.bfloat16, "fp16": torch.float16,
                "bf16x2": torch.bfloat16}
# elements of the torch dtype per stored row: split bf16 keeps hi[1024] | lo[1024]
ROW_ELEMS = {"fp32": nat.SQE_DIM, "bf16": nat.SQE_DIM, "fp16": nat.SQE_DIM, "bf16x2": 2 * nat.SQE_DIM}
ROW_BYTES = {"fp32": 4096, "bf16": 2048, "fp16": 2048, "bf16x2": 4096}
_NAMES = {torch.float32: "fp32", torch.bfloat16: "bf16", torch.float16: "fp16"}

_workspaces = {}
# A call = (memset +) kernels that share one workspace.  Two host threads launching on the same
# stream must not interleave their launches (thread B's memset would land before thread A's
# kernel and B would then see A's leftovers), so every workspace-using call is enqueued under
# this lock.  It only covers the (asynchronous) enqueue, not the execution.
_launch_lock = _threading.Lock()


def dtype_name(t: torch.Tensor) -> str:
    """Storage class of a shard / query tensor: by torch dtype, split bf16 by its 2048-wide rows."""
    try:
        name = _NAMES[t.dtype]
    except KeyError:
        raise TypeError(f"unsupported storage dtype {t.dtype}")
    if name == "bf16" and t.dim() == 2 and t.shape[1] == 2 * nat.SQE_DIM:
        return "bf16x2"
    return name


def _require_cuda(*tensors: torch.Tensor) -> torch.device:
    dev = tensors[0].device
    for t in tensors:
        if not t.is_cuda:
            raise RuntimeError("sqe_b200 has no CPU path: tensors must live on a CUDA device")
        if t.device != dev:
            raise RuntimeError("tensors on different devices")
        if not t.is_contiguous():
            raise ValueError("tensors must be contiguous")
    return dev


def _stream(dev: torch.device) -> int:
    return torch.cuda.current_stream(dev).cuda_stream


def _workspace(dev: torch.device, kind: str, nbytes: int) -> torch.Tensor:
    key = (dev.index, kind, torch.cuda.current_stream(dev).cuda_stream)
    ws = _workspaces.get(key)
    if ws is None or ws.numel() < nbytes:
        ws = torch.zeros(max(int(nbytes), 256), dtype=torch.uint8, device=dev)
        _workspaces[key] = ws
    return ws


def packed_topk_out(dev: torch.device, b: int, k: int):
    """One device buffer holding rows [b,k] int64 then scores [b,k] fp32 (a single D2H copy
    brings both back).  Returns (buffer uint8, scores view, rows view)."""
    buf = torch.empty((b * k * 12,), dtype=torch.uint8, device=dev)
    idx = buf[: b * k * 8].view(torch.int64).view(b, k)
    scores = buf[b * k * 8:].view(torch.float32).view(b, k)
    return buf, scores, idx


def _outputs(out, dev, b, k):
    if out is None:
        return (torch.empty((b, k), dtype=torch.float32, device=dev),
                torch.empty((b, k), dtype=torch.int64, device=dev))
    scores, idx = out
    if scores.shape != (b, k) or idx.shape != (b, k) or scores.dtype != torch.float32 or \
            idx.dtype != torch.int64 or not scores.is_contiguous() or not idx.is_contiguous():
        raise ValueError("bad `out` tensors")
    return scores, idx


def normalize_cast(x: torch.Tensor, dtype: str, out: Optional[torch.Tensor] = None) -> torch.Tensor:
    """K1: rows of fp32 `x` [n,1024] -> x/(||x||+1e-9) stored as `dtype`."""
    dev = _require_cuda(x)
    if x.dtype != torch.float32 or x.dim() != 2 or x.shape[1] != nat.SQE_DIM:
        raise ValueError(f"expected fp32 [n,{nat.SQE_DIM}], got {x.dtype} {tuple(x.shape)}")
    shape = (x.shape[0], ROW_ELEMS[dtype])
    if out is None:
        out = torch.empty(shape, dtype=TORCH_DTYPES[dtype], device=dev)
    else:
        _require_cuda(x, out)
        if tuple(out.shape) != shape or out.dtype != TORCH_DTYPES[dtype]:
            raise ValueError("bad `out`")
    with torch.cuda.device(dev):
        nat.call("sqe_normalize_cast", x.data_ptr(), out.data_ptr(), x.shape[0], x.shape[1],
                 nat.DTYPE_CODES[dtype], _stream(dev))
    return out


def _check_dq(D: torch.Tensor, Q: torch.Tensor, n: Optional[int]) -> Tuple[torch.device, int, int]:
    dev = _require_cuda(D, Q)
    if D.dim() != 2 or Q.dim() != 2 or D.shape[1] not in (nat.SQE_DIM, 2 * nat.SQE_DIM):
        raise ValueError("D and Q must be [rows,1024] (or [rows,2048] bf16 for split bf16)")
    if D.dtype != Q.dtype or D.shape[1] != Q.shape[1] or ROW_ELEMS[dtype_name(D)] != D.shape[1]:
        raise TypeError("queries must be stored in the shard's storage class (use normalize_cast)")
    rows = D.shape[0] if n is None else int(n)
    if rows > D.shape[0]:
        raise ValueError("n exceeds shard rows")
    return dev, rows, Q.shape[0]


def topk_gemv(D: torch.Tensor, Q: torch.Tensor, k: int, idx_offset: int = 0,
              n: Optional[int] = None, out=None) -> Tuple[torch.Tensor, torch.Tensor]:
    """K3: exact cosine top-k, one streaming pass over the shard per query."""
    dev, rows, b = _check_dq(D, Q, n)
    scores, idx = _outputs(out, dev, b, k)
    if b == 0:
        return scores, idx
    with _launch_lock, torch.cuda.device(dev):
        need = nat.load().sqe_topk_gemv_workspace_bytes(b, k)
        ws = _workspace(dev, "gemv", need)
        nat.call("sqe_topk_gemv", D.data_ptr(), nat.DTYPE_CODES[dtype_name(D)], rows, nat.SQE_DIM,
                 Q.data_ptr(), b, k, scores.data_ptr(), idx.data_ptr(), idx_offset,
                 ws.data_ptr(), ws.numel(), _stream(dev))
    return scores, idx


def _xchg_args(xchg, nq: int, k: int):
    """`xchg = (rank, peer pointers, capacity_entries, epoch)` -> the trailing C arguments of the
    sharded entry points (rank, world, peer_buffers_host, capacity_entries, epoch)."""
    if xchg is None:
        return 0, 1, None, 0, 0
    rank, peer_ptrs, cap, epoch = xchg
    world = len(peer_ptrs)
    if nq > nat.SQE_MAX_NQ_FUSED_EXCHANGE:
        raise ValueError(f"the fused exchange takes at most {nat.SQE_MAX_NQ_FUSED_EXCHANGE} queries per call")
    arr = (ctypes.c_void_p * world)(*[int(p) for p in peer_ptrs])
    return int(rank), world, arr, int(cap), int(epoch) & 0xFFFFFFFF


def search_gemv(D: torch.Tensor, q_raw: torch.Tensor, k: int, idx_offset: int = 0,
                n: Optional[int] = None, out=None, ws: Optional[torch.Tensor] = None, xchg=None,
                queries_ready: bool = False) -> Tuple[torch.Tensor, torch.Tensor]:
    """K1 (query side) + K3 in one launch: `q_raw` fp32 [nq,1024] un-normalised.  `ws`: a private
    workspace (zero-initialised uint8 tensor) instead of the per-stream one -- a captured CUDA
    graph must own the memory its kernel uses.  `xchg = (rank, peer pointers, capacity, epoch)`:
    the sharded form -- the query's last CTA also exchanges the rank-local list with the other
    ranks and the outputs are the GLOBAL top-k (`sqe_search_gemv_sharded`, nq <= 2).
    `queries_ready=True`: the caller states that `q_raw` was complete before the previous kernel
    on this stream was launched (a resident tensor, not the output of the kernel just before
    this call); the scan may then overlap the tail of the previous scan (SQE_FLAG_QUERIES_READY)."""
    dev = _require_cuda(D, q_raw)
    if q_raw.dtype != torch.float32 or q_raw.dim() != 2 or q_raw.shape[1] != nat.SQE_DIM:
        raise ValueError("q_raw must be fp32 [nq,1024]")
    if D.dim() != 2 or D.shape[1] != ROW_ELEMS[dtype_name(D)]:
        raise ValueError("D must be [rows,1024] (or [rows,2048] bf16 for split bf16)")
    rows = D.shape[0] if n is None else int(n)
    if rows > D.shape[0]:
        raise ValueError("n exceeds shard rows")
    b = q_raw.shape[0]
    scores, idx = _outputs(out, dev, b, k)
    if b == 0:
        return scores, idx
    with _launch_lock, torch.cuda.device(dev):
        need = nat.load().sqe_topk_gemv_workspace_bytes(b, k)
        if ws is None:
            ws = _workspace(dev, "gemv", need)
        elif ws.numel() < need or ws.device != dev or ws.dtype != torch.uint8:
            raise ValueError("bad private workspace")
        if xchg is None and not queries_ready:
            nat.call("sqe_search_gemv", D.data_ptr(), nat.DTYPE_CODES[dtype_name(D)], rows, nat.SQE_DIM,
                     q_raw.data_ptr(), b, k, scores.data_ptr(), idx.data_ptr(), idx_offset,
                     ws.data_ptr(), ws.numel(), _stream(dev))
        else:
            nat.call("sqe_search_gemv_sharded", D.data_ptr(), nat.DTYPE_CODES[dtype_name(D)], rows, nat.SQE_DIM,
                     q_raw.data_ptr(), b, k, scores.data_ptr(), idx.data_ptr(), idx_offset,
                     *_xchg_args(xchg, b, k), nat.SQE_FLAG_QUERIES_READY if queries_ready else 0,
                     ws.data_ptr(), ws.numel(), _stream(dev))
    return scores, idx


def search_gemv_prefiltered(D: torch.Tensor, D8: torch.Tensor, meta: torch.Tensor, q_raw: torch.Tensor,
                            k: int, idx_offset: int = 0, n: Optional[int] = None, out=None,
                            rescored: Optional[torch.Tensor] = None, ws: Optional[torch.Tensor] = None,
                            xchg=None, queries_ready: bool = False) -> Tuple[torch.Tensor, torch.Tensor]:
    """K1 (query side) + K3p in two launches (no separate normalise launch): `q_raw` fp32 [nq,1024]
    un-normalised; both passes normalise the query themselves.  Bit-identical to `normalize_cast`
    + `topk_gemv_prefiltered`.  `xchg`: as in `search_gemv` (the exchange rides in the rescoring
    pass's last CTA)."""
    dev = _require_cuda(D, D8, meta, q_raw)
    if q_raw.dtype != torch.float32 or q_raw.dim() != 2 or q_raw.shape[1] != nat.SQE_DIM:
        raise ValueError("q_raw must be fp32 [nq,1024]")
    if D.dim() != 2 or D.shape[1] != ROW_ELEMS[dtype_name(D)]:
        raise ValueError("D must be [rows,1024] (or [rows,2048] bf16 for split bf16)")
    rows = D.shape[0] if n is None else int(n)
    if rows > D.shape[0]:
        raise ValueError("n exceeds shard rows")
    if D8.dtype != torch.int8 or meta.dtype != torch.float32 or D8.dim() != 2 or meta.dim() != 2 or \
            D8.shape[1] != nat.SQE_DIM or meta.shape[1] != 4 or D8.shape[0] < rows or meta.shape[0] < rows:
        raise ValueError("coarse rows must be int8 [rows,1024] + fp32 [rows,4] (ops.quantize_rows)")
    b = q_raw.shape[0]
    if b > nat.SQE_MAX_NQ_PREFILTER:
        raise ValueError(f"at most {nat.SQE_MAX_NQ_PREFILTER} queries per call")
    scores, idx = _outputs(out, dev, b, k)
    if b == 0:
        return scores, idx
    if rescored is not None and (rescored.dtype != torch.int32 or rescored.numel() < b or not rescored.is_cuda):
        raise ValueError("`rescored` must be an int32 CUDA tensor [nq]")
    with _launch_lock, torch.cuda.device(dev):
        need = nat.load().sqe_topk_gemv_prefiltered_workspace_bytes(rows, b, k)
        if ws is None:
            ws = _workspace(dev, "prefilter", need)
        elif ws.numel() < need or ws.device != dev or ws.dtype != torch.uint8:
            raise ValueError("bad private workspace")
        nat.call("sqe_search_gemv_prefiltered", D.data_ptr(), nat.DTYPE_CODES[dtype_name(D)], rows,
                 nat.SQE_DIM, D8.data_ptr(), meta.data_ptr(), q_raw.data_ptr(), b, k, scores.data_ptr(),
                 idx.data_ptr(), idx_offset, rescored.data_ptr() if rescored is not None else None,
                 *_xchg_args(xchg, b, k), nat.SQE_FLAG_QUERIES_READY if queries_ready else 0,
                 ws.data_ptr(), ws.numel(), _stream(dev))
    return scores, idx


def quantize_rows(D: torch.Tensor, n: Optional[int] = None, out=None, row0: int = 0
                  ) -> Tuple[torch.Tensor, torch.Tensor]:
    """K1q: the coarse copy of stored rows for the prefiltered scan: (int8 [n,1024], fp32 [n,4] =
    {scale, |d - scale d8| bound, |scale d8| bound, 0}).  `out=(D8, meta)` + `row0`: write rows
    [row0, row0+n) of preallocated buffers (ingest into a shard tail)."""
    dev = _require_cuda(D)
    if D.dim() != 2 or D.shape[1] != ROW_ELEMS[dtype_name(D)]:
        raise ValueError("D must be [rows,1024] (or [rows,2048] bf16 for split bf16)")
    rows = D.shape[0] if n is None else int(n)
    if rows > D.shape[0]:
        raise ValueError("n exceeds shard rows")
    if out is None:
        d8 = torch.empty((rows, nat.SQE_DIM), dtype=torch.int8, device=dev)
        meta = torch.empty((rows, 4), dtype=torch.float32, device=dev)
        row0 = 0
    else:
        d8, meta = out
        _require_cuda(D, d8, meta)
        if d8.dtype != torch.int8 or meta.dtype != torch.float32 or d8.dim() != 2 or meta.dim() != 2 or \
                d8.shape[1] != nat.SQE_DIM or meta.shape[1] != 4 or \
                d8.shape[0] < row0 + rows or meta.shape[0] < row0 + rows:
            raise ValueError("bad `out` buffers")
    if rows:
        with torch.cuda.device(dev):
            nat.call("sqe_quantize_rows", D.data_ptr(), nat.DTYPE_CODES[dtype_name(D)], rows, nat.SQE_DIM,
                     d8.data_ptr() + row0 * nat.SQE_DIM, meta.data_ptr() + row0 * 16, _stream(dev))
    return d8, meta


def topk_gemv_prefiltered(D: torch.Tensor, D8: torch.Tensor, meta: torch.Tensor, Q: torch.Tensor,
                          k: int, idx_offset: int = 0, n: Optional[int] = None, out=None,
                          rescored: Optional[torch.Tensor] = None, ws: Optional[torch.Tensor] = None
                          ) -> Tuple[torch.Tensor, torch.Tensor]:
    """K3p: the results of `topk_gemv`, bit for bit, at about half the HBM traffic: int8 scan of
    the coarse copy (`quantize_rows`) with a rigorous error bound, then exact rescoring of the
    rows the bound cannot rule out.  `rescored`: optional int32 CUDA tensor [nq] that receives the
    number of rows rescored per query.  `ws`: a private zero-initialised workspace (CUDA graphs)."""
    dev, rows, b = _check_dq(D, Q, n)
    _require_cuda(D, D8, meta)
    if D8.dtype != torch.int8 or meta.dtype != torch.float32 or D8.dim() != 2 or meta.dim() != 2 or \
            D8.shape[1] != nat.SQE_DIM or meta.shape[1] != 4 or D8.shape[0] < rows or meta.shape[0] < rows:
        raise ValueError("coarse rows must be int8 [rows,1024] + fp32 [rows,4] (ops.quantize_rows)")
    if b > nat.SQE_MAX_NQ_PREFILTER:
        raise ValueError(f"at most {nat.SQE_MAX_NQ_PREFILTER} queries per call")
    scores, idx = _outputs(out, dev, b, k)
    if b == 0:
        return scores, idx
    if rescored is not None and (rescored.dtype != torch.int32 or rescored.numel() < b or not rescored.is_cuda):
        raise ValueError("`rescored` must be an int32 CUDA tensor [nq]")
    with _launch_lock, torch.cuda.device(dev):
        need = nat.load().sqe_topk_gemv_prefiltered_workspace_bytes(rows, b, k)
        if ws is None:
            ws = _workspace(dev, "prefilter", need)
        elif ws.numel() < need or ws.device != dev or ws.dtype != torch.uint8:
            raise ValueError("bad private workspace")
        nat.call("sqe_topk_gemv_prefiltered", D.data_ptr(), nat.DTYPE_CODES[dtype_name(D)], rows,
                 nat.SQE_DIM, D8.data_ptr(), meta.data_ptr(), Q.data_ptr(), b, k, scores.data_ptr(),
                 idx.data_ptr(), idx_offset, rescored.data_ptr() if rescored is not None else None,
                 ws.data_ptr(), ws.numel(), _stream(dev))
    return scores, idx


def search_batched_prefiltered(D: torch.Tensor, D8: torch.Tensor, meta: torch.Tensor, q_raw: torch.Tensor,
                               k: int, idx_offset: int = 0, n: Optional[int] = None, out=None,
                               rescored: Optional[torch.Tensor] = None) -> Tuple[torch.Tensor, torch.Tensor]:
    """K2p: exact cosine top-k for a batch of RAW fp32 queries from the int8 copy on the tensor cores
    (`quantize_rows`) + exact rescoring: the results of `normalize_cast` + `topk_gemv`, bit for
    bit, for any storage class of `D` (fp32 shards included).  `rescored`: optional int32 CUDA
    tensor [b] that receives the number of rows scored exactly per query."""
    dev = _require_cuda(D, D8, meta, q_raw)
    if q_raw.dtype != torch.float32 or q_raw.dim() != 2 or q_raw.shape[1] != nat.SQE_DIM:
        raise ValueError("q_raw must be fp32 [b,1024]")
    if D.dim() != 2 or D.shape[1] != ROW_ELEMS[dtype_name(D)]:
        raise ValueError("D must be [rows,1024] (or [rows,2048] bf16 for split bf16)")
    rows = D.shape[0] if n is None else int(n)
    if rows > D.shape[0]:
        raise ValueError("n exceeds shard rows")
    if D8.dtype != torch.int8 or meta.dtype != torch.float32 or D8.dim() != 2 or meta.dim() != 2 or \
            D8.shape[1] != nat.SQE_DIM or meta.shape[1] != 4 or D8.shape[0] < rows or meta.shape[0] < rows:
        raise ValueError("coarse rows must be int8 [rows,1024] + fp32 [rows,4] (ops.quantize_rows)")
    b = q_raw.shape[0]
    scores, idx = _outputs(out, dev, b, k)
    if b == 0:
        return scores, idx
    if rescored is not None and (rescored.dtype != torch.int32 or rescored.numel() < b or not rescored.is_cuda):
        raise ValueError("`rescored` must be an int32 CUDA tensor [b]")
    code = nat.DTYPE_CODES[dtype_name(D)]
    with _launch_lock, torch.cuda.device(dev):
        need = nat.load().sqe_search_batched_prefiltered_workspace_bytes(rows, b, k, code)
        ws = _workspace(dev, "batched_i8", need)
        nat.call("sqe_search_batched_prefiltered", D.data_ptr(), code, rows, nat.SQE_DIM, D8.data_ptr(),
                 meta.data_ptr(), q_raw.data_ptr(), b, k, scores.data_ptr(), idx.data_ptr(), idx_offset,
                 rescored.data_ptr() if rescored is not None else None, ws.data_ptr(), ws.numel(), _stream(dev))
    return scores, idx


def k2p_pays(rows: int, b: int, dtype: str) -> bool:
    """Routing rule between the two tensor-core batch paths, from the measured grid
    (profiles/r2_k2p_grid.txt, 0.5M-10M rows x b = 32..1024): K2p's scan moves half the bytes at
    twice the tensor rate but pays ~0.1-0.3 ms of fixed work (query preparation, the bound's
    start-up phase, the exact pass), so it wins from ~1M rows for small batches and from ~2.5M
    rows for b = 1024 (1.3-1.7 x at 4M rows, 1.5-1.6 x at 10M).  fp32 shards: always (their only
    other batch path is one streaming pass per query)."""
    if dtype == "fp32":
        return True
    need = 1_000_000 if b <= 64 else 1_500_000 if b <= 128 else 2_000_000 if b <= 512 else 2_500_000
    return rows >= need


def topk_batched(D: torch.Tensor, Q: torch.Tensor, k: int, idx_offset: int = 0,
                 n: Optional[int] = None, out=None) -> Tuple[torch.Tensor, torch.Tensor]:
    """K2: exact cosine top-k on the tensor cores (bf16/fp16 shards)."""
    dev, rows, b = _check_dq(D, Q, n)
    scores, idx = _outputs(out, dev, b, k)
    if b == 0:
        return scores, idx
    with _launch_lock, torch.cuda.device(dev):
        need = nat.load().sqe_topk_batched_workspace_bytes(rows, b, k)
        ws = _workspace(dev, "batched", need)
        nat.call("sqe_topk_batched", D.data_ptr(), nat.DTYPE_CODES[dtype_name(D)], rows,
                 nat.SQE_DIM, Q.data_ptr(), b, k, scores.data_ptr(), idx.data_ptr(), idx_offset,
                 ws.data_ptr(), ws.numel(), _stream(dev))
    return scores, idx


def topk(D: torch.Tensor, Q: torch.Tensor, k: int, idx_offset: int = 0,
         n: Optional[int] = None, out=None) -> Tuple[torch.Tensor, torch.Tensor]:
    """Route: one query or an fp32 shard -> K3 (HBM-bound GEMV); else K2 (tensor cores)."""
    if Q.shape[0] <= 1 or D.dtype == torch.float32 or k > nat.SQE_MAX_K_BATCHED:
        return topk_gemv(D, Q, k, idx_offset, n, out=out)
    return topk_batched(D, Q, k, idx_offset, n, out=out)


def packed_cache_out(dev: torch.device, b: int):
    """One device buffer for (idx int32 [b], score fp32 [b], hit uint8 [b]): a single D2H copy."""
    buf = torch.empty((b * 9,), dtype=torch.uint8, device=dev)
    idx = buf[: b * 4].view(torch.int32)
    score = buf[b * 4: b * 8].view(torch.float32)
    hit = buf[b * 8:]
    return buf, idx, score, hit


def cache_top1(C: torch.Tensor, Q: torch.Tensor, threshold: float, path: int = 0,
               n: Optional[int] = None, out=None) -> Tuple[torch.Tensor, torch.Tensor, torch.Tensor]:
    """K5: (idx int32 [b], score fp32 [b], hit uint8 [b])."""
    dev, rows, b = _check_dq(C, Q, n)
    if out is None:
        score = torch.empty((b,), dtype=torch.float32, device=dev)
        idx = torch.empty((b,), dtype=torch.int32, device=dev)
        hit = torch.empty((b,), dtype=torch.uint8, device=dev)
    else:
        idx, score, hit = out
    if b == 0:
        return idx, score, hit
    with _launch_lock, torch.cuda.device(dev):
        need = nat.load().sqe_cache_top1_workspace_bytes(rows, b)
        ws = _workspace(dev, "cache", need)
        nat.call("sqe_cache_top1", C.data_ptr(), nat.DTYPE_CODES[dtype_name(C)], rows, nat.SQE_DIM,
                 Q.data_ptr(), b, float(threshold), score.data_ptr(), idx.data_ptr(),
                 hit.data_ptr(), path, ws.data_ptr(), ws.numel(), _stream(dev))
        if path == 2 or (path == 0 and C.dtype != torch.float32 and b > 1):
            nat.launch_count += 1
    return idx, score, hit


def cache_top1_prefiltered(C: torch.Tensor, C8: torch.Tensor, meta: torch.Tensor, q_raw: torch.Tensor,
                           threshold: float, n: Optional[int] = None, out=None
                           ) -> Tuple[torch.Tensor, torch.Tensor, torch.Tensor]:
    """K5 through K2p: (idx int32 [b], score fp32 [b], hit uint8 [b]) for RAW fp32 queries, from the
    int8 copy of the cache rows + exact rescoring (any storage class of `C`)."""
    dev = _require_cuda(C, C8, meta, q_raw)
    if q_raw.dtype != torch.float32 or q_raw.dim() != 2 or q_raw.shape[1] != nat.SQE_DIM:
        raise ValueError("q_raw must be fp32 [b,1024]")
    rows = C.shape[0] if n is None else int(n)
    if rows > C.shape[0] or C8.shape[0] < rows or meta.shape[0] < rows:
        raise ValueError("n exceeds the cache rows / their coarse copy")
    b = q_raw.shape[0]
    if out is None:
        score = torch.empty((b,), dtype=torch.float32, device=dev)
        idx = torch.empty((b,), dtype=torch.int32, device=dev)
        hit = torch.empty((b,), dtype=torch.uint8, device=dev)
    else:
        idx, score, hit = out
    if b == 0:
        return idx, score, hit
    code = nat.DTYPE_CODES[dtype_name(C)]
    with _launch_lock, torch.cuda.device(dev):
        need = nat.load().sqe_cache_top1_prefiltered_workspace_bytes(rows, b, code)
        ws = _workspace(dev, "cache_i8", need)
        nat.call("sqe_cache_top1_prefiltered", C.data_ptr(), code, rows, nat.SQE_DIM, C8.data_ptr(), meta.data_ptr(),
                 q_raw.data_ptr(), b, float(threshold), score.data_ptr(), idx.data_ptr(), hit.data_ptr(),
                 ws.data_ptr(), ws.numel(), _stream(dev))
    return idx, score, hit


def merge_topk(scores: torch.Tensor, idx: torch.Tensor, k_out: int, out=None
               ) -> Tuple[torch.Tensor, torch.Tensor]:
    """K4: scores/idx [lists,b,k_in] -> best-first [b,k_out]."""
    dev = _require_cuda(scores, idx)
    if scores.dim() != 3 or scores.shape != idx.shape:
        raise ValueError("scores/idx must be [lists,b,k]")
    if scores.dtype != torch.float32 or idx.dtype != torch.int64:
        raise TypeError("scores fp32, idx int64")
    lists, b, k_in = scores.shape
    out_s, out_i = _outputs(out, dev, b, k_out)
    if b == 0:
        return out_s, out_i
    with torch.cuda.device(dev):
        nat.call("sqe_merge_topk", scores.data_ptr(), idx.data_ptr(), lists, b, k_in, k_out,
                 out_s.data_ptr(), out_i.data_ptr(), _stream(dev))
    return out_s, out_i


def exchange_buffer_bytes(world: int, capacity_entries: int) -> int:
    n = nat.load().sqe_exchange_buffer_bytes(world, capacity_entries)
    if n < 0:
        raise ValueError("bad exchange buffer size arguments")
    return int(n)


def exchange_merge(scores: torch.Tensor, idx: torch.Tensor, k_out: int, rank: int,
                   peer_ptrs, capacity_entries: int, epoch: int, wait_mask: Optional[int] = None,
                   out=None) -> Tuple[torch.Tensor, torch.Tensor]:
    """K4x: push this rank's [b,k] lists into every rank's peer-mapped buffer, wait for all
    ranks, merge.  `peer_ptrs`: device pointers (ints) of every rank's buffer as mapped here."""
    dev = _require_cuda(scores, idx)
    if scores.dim() != 2 or scores.shape != idx.shape:
        raise ValueError("scores/idx must be [b,k]")
    if scores.dtype != torch.float32 or idx.dtype != torch.int64:
        raise TypeError("scores fp32, idx int64")
    b, k_in = scores.shape
    world = len(peer_ptrs)
    out_s, out_i = _outputs(out, dev, b, k_out)
    if b == 0:
        return out_s, out_i
    arr = (ctypes.c_void_p * world)(*[int(p) for p in peer_ptrs])
    if wait_mask is None:
        wait_mask = (1 << world) - 1
    with torch.cuda.device(dev):
        nat.call("sqe_exchange_merge", scores.data_ptr(), idx.data_ptr(), b, k_in, k_out, rank, world,
                 arr, capacity_entries, epoch & 0xFFFFFFFF, wait_mask, out_s.data_ptr(),
                 out_i.data_ptr(), _stream(dev))
    return out_s, out_i


class SingleQueryGraph:
    """A captured CUDA graph of the launch-bound single-query step:
    pinned host query -> device, `sqe_search_gemv` (normalise + scan + top-k), device -> pinned host.
    Replaying it costs one graph launch instead of several Python-level copies and launches.
    The shard pointer, row count, k and row offset are baked in: build a new one when they change."""

    def __init__(self, shard: torch.Tensor, rows: int, k: int, idx_offset: int = 0, coarse=None):
        """`coarse = (D8, meta, dtype name)`: capture the prefiltered scan (K3p, raw-query form) instead of the
        fused exact scan; same results, about half the bytes per replay."""
        dev = shard.device
        self.coarse = coarse
        self.device = dev
        self.k = int(k)
        self.idx_offset = int(idx_offset)
        self.rows = int(rows)
        self.shard_ptr = shard.data_ptr()
        self._shard = shard                                    # keep the storage alive
        self.host_q = torch.empty((1, nat.SQE_DIM), dtype=torch.float32).pin_memory()
        self.host_out = torch.empty((self.k * 12,), dtype=torch.uint8).pin_memory()
        self.dev_q = torch.empty((1, nat.SQE_DIM), dtype=torch.float32, device=dev)
        self.dev_out, self.scores, self.idx = packed_topk_out(dev, 1, self.k)
        # private workspace: the per-stream ones are shared by everything launched on that stream
        # and replaced when they grow, neither of which a captured pointer survives
        if coarse is None:
            need = nat.load().sqe_topk_gemv_workspace_bytes(1, self.k)
        else:
            need = nat.load().sqe_topk_gemv_prefiltered_workspace_bytes(self.rows, 1, self.k)
        self._ws = torch.zeros((int(need),), dtype=torch.uint8, device=dev)
        self.host_q.zero_()
        with torch.cuda.device(dev):
            side = torch.cuda.Stream(device=dev)
            side.wait_stream(torch.cuda.current_stream(dev))
            with torch.cuda.stream(side):                      # warm-up on the side stream (allocates its workspace)
                self._body()
            side.synchronize()
            self.graph = torch.cuda.CUDAGraph()
            # thread_local: CUDA calls of OTHER host threads (an ingest thread's cudaMalloc / pageable
            # copy / stream sync, main.py:454-455) neither fail nor invalidate this capture
            with torch.cuda.graph(self.graph, stream=side, capture_error_mode="thread_local"):
                self._body()
        self._np_q = self.host_q.numpy()
        self._np_out = self.host_out.numpy()

    def _body(self) -> None:
        self.dev_q.copy_(self.host_q, non_blocking=True)
        if self.coarse is not None:
            search_gemv_prefiltered(self._shard, self.coarse[0], self.coarse[1], self.dev_q, self.k,
                                    idx_offset=self.idx_offset, n=self.rows, out=(self.scores, self.idx), ws=self._ws)
        else:
            search_gemv(self._shard, self.dev_q, self.k, idx_offset=self.idx_offset, n=self.rows,
                        out=(self.scores, self.idx), ws=self._ws)
        self.host_out.copy_(self.dev_out, non_blocking=True)

    def run(self, q_row: np.ndarray) -> Tuple[np.ndarray, np.ndarray]:
        """q_row: fp32 [1024] (raw).  Returns (scores [k] fp32, rows [k] int64) as fresh arrays."""
        self._np_q[0, :] = q_row
        self.graph.replay()
        nat.launch_count += 1 if self.coarse is None else 2    # the replayed sqe_search_gemv | K3p (two passes)
        torch.cuda.current_stream(self.device).synchronize()
        k = self.k
        return self._np_out[k * 8:].view(np.float32).copy(), self._np_out[: k * 8].view(np.int64).copy()


class _StreamSlot:
    """Buffers and events of one in-flight batch of `stream_pipeline` (reused round robin)."""

    def __init__(self):
        self.host_q: Optional[torch.Tensor] = None      # pinned staging for pageable inputs
        self.dev_q: Optional[torch.Tensor] = None
        self.dev_out: Optional[torch.Tensor] = None     # packed result bytes
        self.host_out: Optional[torch.Tensor] = None    # pinned
        self.ev_h2d = torch.cuda.Event()
        self.ev_done = torch.cuda.Event()
        self.ev_d2h = torch.cuda.Event()

    def stage(self, q: np.ndarray, b: int, out_bytes: int, dev: torch.device) -> torch.Tensor:
        """Size the buffers for a [b,1024] batch; returns the page-locked source of the
        host->device copy (the caller's own array when it is already pinned)."""
        if self.dev_q is None or self.dev_q.shape[0] < b:
            self.dev_q = torch.empty((b, nat.SQE_DIM), dtype=torch.float32, device=dev)
        if self.dev_out is None or self.dev_out.numel() < out_bytes:
            self.dev_out = torch.empty((out_bytes,), dtype=torch.uint8, device=dev)
            self.host_out = torch.empty((out_bytes,), dtype=torch.uint8).pin_memory()
        if q is None:                                   # the queries are produced on the device
            return None
        t = torch.from_numpy(q)
        if t.is_pinned():
            return t
        if self.host_q is None or self.host_q.shape[0] < b:
            self.host_q = torch.empty((b, nat.SQE_DIM), dtype=torch.float32).pin_memory()
        self.host_q[:b].copy_(t)
        return self.host_q[:b]


class StreamPipeline:
    """Copy/compute overlap for a stream of host query batches (behind
    `GpuCorpusIndex.search_batches`, `GpuQueryCache.lookup_batches` and the MicroBatcher).

    `submit(batch, ctx)`: `as_rows(batch)` -> fp32 [b,1024]; its host->device copy runs on a copy
    stream, `launch(q_dev [b,1024], packed_out uint8 [out_bytes(b, ctx)], ctx)` enqueues the
    kernels on the compute stream (the constructing thread's current stream), the packed result
    returns on a second copy stream.  `collect()` waits for the OLDEST in-flight batch and returns
    `unpack(host uint8 array, b, ctx)` (fresh host arrays).  At most `depth` batches are in flight;
    `submit` blocks while all slots are taken, so a second thread may do the collecting."""

    def __init__(self, dev: torch.device, as_rows, out_bytes, launch, unpack, depth: int = 2):
        if depth < 1:
            raise ValueError("depth must be >= 1")
        self.dev, self.depth = dev, depth
        self._as_rows, self._out_bytes, self._launch, self._unpack = as_rows, out_bytes, launch, unpack
        with torch.cuda.device(dev):
            self.compute = torch.cuda.current_stream(dev)
            self.h2d, self.d2h = torch.cuda.Stream(device=dev), torch.cuda.Stream(device=dev)
        self._free = queue.Queue()
        for _ in range(depth):
            self._free.put(_StreamSlot())
        self._pending = queue.Queue()

    def in_flight(self) -> int:
        return self._pending.qsize()

    def submit(self, batch, ctx=None) -> None:
        q = self._as_rows(batch)
        b = q.shape[0]
        nbytes = self._out_bytes(b, ctx)
        s = self._free.get()
        try:
            with torch.cuda.device(self.dev), torch.cuda.stream(self.compute):
                src = s.stage(q, b, nbytes, self.dev)
                with torch.cuda.stream(self.h2d):
                    s.dev_q[:b].copy_(src, non_blocking=True)
                    s.ev_h2d.record(self.h2d)
                self.compute.wait_event(s.ev_h2d)
                self._launch(s.dev_q[:b], s.dev_out[:nbytes], ctx)
                s.ev_done.record(self.compute)
                with torch.cuda.stream(self.d2h):
                    self.d2h.wait_event(s.ev_done)
                    s.host_out[:nbytes].copy_(s.dev_out[:nbytes], non_blocking=True)
                    s.ev_d2h.record(self.d2h)
        except BaseException:
            self._free.put(s)
            raise
        self._pending.put((s, b, nbytes, ctx))

    def submit_device(self, make_q, b: int, ctx=None) -> None:
        """Like `submit`, for queries that are PRODUCED on the device: `make_q()` enqueues its work
        on the compute stream and returns q_dev [b,1024] fp32 (e.g. the embedding encoder's forward
        pass for a batch of texts) -- no host rows, no H2D stage."""
        nbytes = self._out_bytes(b, ctx)
        s = self._free.get()
        try:
            with torch.cuda.device(self.dev), torch.cuda.stream(self.compute):
                s.stage(None, b, nbytes, self.dev)
                qd = make_q()
                self._launch(qd, s.dev_out[:nbytes], ctx)
                s.ev_done.record(self.compute)
                with torch.cuda.stream(self.d2h):
                    self.d2h.wait_event(s.ev_done)
                    s.host_out[:nbytes].copy_(s.dev_out[:nbytes], non_blocking=True)
                    s.ev_d2h.record(self.d2h)
        except BaseException:
            self._free.put(s)
            raise
        self._pending.put((s, b, nbytes, ctx))

    def collect(self):
        s, b, nbytes, ctx = self._pending.get()
        try:
            s.ev_d2h.synchronize()
            return self._unpack(s.host_out[:nbytes].numpy(), b, ctx)
        finally:
            self._free.put(s)


def stream_pipeline(dev: torch.device, batches, as_rows, out_bytes, launch, unpack, depth: int = 2):
    """Single-threaded generator over a StreamPipeline: the result of batch j-depth is yielded when
    batch j has been submitted, so the copies of the neighbouring batches overlap the kernels of
    the current one.  The callbacks take no ctx here."""
    pipe = StreamPipeline(dev, as_rows, lambda b, _c: out_bytes(b), lambda qd, buf, _c: launch(qd, buf),
                          lambda arr, b, _c: unpack(arr, b), depth)
    for batch in batches:
        done = pipe.collect() if pipe.in_flight() == depth else None      # frees a slot
        pipe.submit(batch)
        if done is not None:
            yield done
    while pipe.in_flight():
        yield pipe.collect()
