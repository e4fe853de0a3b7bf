"""GPU-resident corpus index: the drop-in for the reference's `OpenSearchIndexer`
(/root/reference/app/main.py:291-373) and `bulk_index_embeddings`
(app/embedding_gen.py:196-257).

Same method names, argument meaning and error behaviour:
  * `add_embeddings(embeddings, docs)`  -- main.py:309-338: rows are L2-normalised with
    `x/(||x||+1e-9)` (K1, on the GPU) and appended; the payload (`doc_id`, `text`) and
    the `_id = f"{doc_id}_{i}"` of main.py:325 are kept on the host;
  * `search(query_emb, k=3)`            -- main.py:347-373: normalise the query, score it
    against every stored row (exact, not HNSW), return `[(source_dict, float_score)]`
    best-first; `[]` on an empty query or on any backend error;
  * `has_any_data()`                    -- main.py:300-307.

What differs, by design: scoring is exact (the reference's external index is approximate and
unpinned), rows live in HBM as bf16/fp16/fp32, and a batched entry point
(`search_batch`) exists for the throughput configurations.
"""
from __future__ import annotations

import json
import os
import threading
import time
from typing import Dict, List, Optional, Sequence, Tuple

import numpy as np
import torch

from . import _native as nat
from . import ops

EMBED_DIM = nat.SQE_DIM          # main.py:38


class GpuCorpusIndex:
    _GRAPH_RETRY_S = 5.0
    K2P_MAX_K = 32          # largest k a batch on a 16-bit shard is answered through the int8 prefilter

    def __init__(self, client=None, index_name: str = "", *, dtype: str = "bf16",
                 device: Optional[torch.device] = None, initial_capacity: int = 65536,
                 score_mode: str = "cosine", strict: bool = False,
                 return_embedding: bool = False, keep_payload: bool = True,
                 use_graphs: bool = True, prefilter: bool = False):
        """`client` / `index_name` are accepted for signature compatibility
        (main.py:296-298) and ignored: there is no OpenSearch behind this index.

        score_mode: "cosine" returns the cosine; "opensearch" returns 1/(2-cos), the
        `_score` an OpenSearch `cosinesimil` index reports.
        strict: raise instead of the reference's print-and-return-[] on errors.
        prefilter: keep an int8 copy of every row (+50 % memory for a bf16 shard) and answer
        through it: one- and two-query searches with the prefiltered scan (K3p), batches with the
        int8 tensor-core prefilter (K2p) -- the SAME results as the exact scan (K3), bit for bit,
        at about half the HBM traffic per query and twice the tensor rate per batch."""
        if dtype not in ops.TORCH_DTYPES:
            raise ValueError(f"dtype must be one of {sorted(ops.TORCH_DTYPES)}")
        if score_mode not in ("cosine", "opensearch"):
            raise ValueError("score_mode must be 'cosine' or 'opensearch'")
        self.client = client
        self.index_name = index_name
        self.dtype = dtype
        self.score_mode = score_mode
        self.strict = strict
        self.return_embedding = return_embedding
        self.keep_payload = keep_payload
        self.device = torch.device(device) if device is not None else torch.device("cuda", 0)
        self._lock = threading.Lock()           # add_embeddings runs in an executor (main.py:454-455)
        self._search_lock = threading.Lock()    # the pinned staging buffers are per index
        self._rows = 0                          # published row count
        self._capacity = 0
        self._shard: Optional[torch.Tensor] = None
        self.prefilter = bool(prefilter)
        self._coarse8: Optional[torch.Tensor] = None     # int8 [capacity,1024]  (prefilter only)
        self._coarse_meta: Optional[torch.Tensor] = None  # fp32 [capacity,4]
        self._initial_capacity = int(initial_capacity)
        self._docs: List[Dict[str, str]] = []   # payload table, row-aligned
        self._ids: List[str] = []
        self._row_of_id: Dict[str, int] = {}    # `_id` -> row: an "index" action with a known _id overwrites
        # CUDA graphs of the single-query step, keyed by k; dropped whenever rows are added
        self.use_graphs = use_graphs
        self._graphs: Dict[int, "ops.SingleQueryGraph"] = {}
        self._graph_seen: Dict[int, tuple] = {}   # index state at the last eager single-query search, per k
        self._graph_retry_at = 0.0               # monotonic time before which no capture is attempted
        self._pinned_q: Optional[torch.Tensor] = None
        self._pinned_out: Optional[torch.Tensor] = None

    # ------------------------------------------------------------------ storage
    @property
    def num_rows(self) -> int:
        return self._rows

    @property
    def shard(self) -> torch.Tensor:
        """The stored rows [num_rows,1024] (a view of the HBM shard)."""
        if self._shard is None:
            return torch.empty((0, ops.ROW_ELEMS[self.dtype]), dtype=ops.TORCH_DTYPES[self.dtype], device=self.device)
        return self._shard[: self._rows]

    def reserve(self, rows: int) -> None:
        """Make room for `rows` rows in total (one allocation, no later regrowth)."""
        with self._lock:
            self._grow_locked(rows)

    def _grow_locked(self, need: int) -> None:
        if need <= self._capacity:
            return
        cap = max(need, self._initial_capacity, self._capacity * 2)
        new = torch.empty((cap, ops.ROW_ELEMS[self.dtype]), dtype=ops.TORCH_DTYPES[self.dtype], device=self.device)
        if self._shard is not None and self._rows:
            new[: self._rows].copy_(self._shard[: self._rows])
        c8 = cm = None
        if self.prefilter:
            c8 = torch.empty((cap, EMBED_DIM), dtype=torch.int8, device=self.device)
            cm = torch.empty((cap, 4), dtype=torch.float32, device=self.device)
            if self._coarse8 is not None and self._rows:
                c8[: self._rows].copy_(self._coarse8[: self._rows])
                cm[: self._rows].copy_(self._coarse_meta[: self._rows])
        # Searches read `_shard` / `_coarse*` without the ingest lock and may run on other streams (a
        # MicroBatcher's compute stream): the new buffers become visible only once the copies into them
        # have finished, and the old ones are released only after every kernel that may still read them
        # has (a search that picked up the old pointers just before the swap).
        old = (self._shard, self._coarse8, self._coarse_meta)
        if old[0] is not None:
            torch.cuda.current_stream(self.device).synchronize()
        self._shard = new
        if self.prefilter:
            self._coarse8, self._coarse_meta = c8, cm
        self._capacity = cap
        if old[0] is not None:
            torch.cuda.synchronize(self.device)
        del old

    def enable_prefilter(self) -> None:
        """Turn the prefiltered scan on for an index that was built without it: allocate the int8
        copy for the current capacity and quantise the rows already stored (K1q, one pass)."""
        with self._lock:
            if self.prefilter and (self._coarse8 is not None or self._shard is None):
                return
            if self._shard is not None:
                c8 = torch.empty((self._capacity, EMBED_DIM), dtype=torch.int8, device=self.device)
                cm = torch.empty((self._capacity, 4), dtype=torch.float32, device=self.device)
                with torch.cuda.device(self.device):
                    if self._rows:
                        ops.quantize_rows(self._shard[: self._rows], out=(c8, cm))
                    torch.cuda.current_stream(self.device).synchronize()
                self._coarse8, self._coarse_meta = c8, cm
            self.prefilter = True                        # last: searches may run concurrently
            self._graphs.clear()

    def has_any_data(self) -> bool:                      # main.py:300-307
        try:
            return self._rows > 0
        except Exception:
            return False

    # ------------------------------------------------------------------- ingest
    def add_embeddings(self, embeddings: np.ndarray, docs: Sequence[Dict[str, str]]) -> None:
        """main.py:309-338.  `embeddings` [N,1024] fp32 (un-normalised), `docs` N dicts with
        "doc_id" and "text"."""
        if embeddings is None or getattr(embeddings, "size", 0) == 0:
            print("[GpuCorpusIndex] No embeddings.")     # main.py:310-312
            return
        try:
            self._add(embeddings, docs, id_from_global_row=True)
        except Exception as e:                           # main.py:344-345 prints and continues
            if self.strict:
                raise
            print(f"[GpuCorpusIndex] Bulk indexing error: {e}")

    def add_document_chunks(self, doc_id: str, embeddings: np.ndarray, chunks: Sequence[str]) -> None:
        """embedding_gen.py:196-257 (`bulk_index_embeddings`): all chunks share `doc_id`,
        `_id = f"{doc_id}_{chunk_index}"`."""
        if embeddings is None or getattr(embeddings, "size", 0) == 0:
            print("[ERROR] Missing embeddings => cannot index.")
            return
        docs = [{"doc_id": doc_id, "text": c} for c in chunks]
        try:
            self._add(embeddings, docs, id_from_global_row=False)
        except Exception as e:
            if self.strict:
                raise
            print(f"[GpuCorpusIndex] Bulk error (doc_id={doc_id}): {e}")

    def _add(self, embeddings, docs, id_from_global_row: bool, ids: Optional[Sequence[str]] = None) -> None:
        """Append rows -- or overwrite them: the reference sends `_op_type: "index"` actions with
        `_id = f"{doc_id}_{i}"` (main.py:321-325, embedding_gen.py:221-233), and OpenSearch REPLACES
        a document whose `_id` already exists (re-uploading a document replaces its chunks instead
        of duplicating them).  A row whose `_id` is known is therefore rewritten in place (it keeps
        its row number, so the tie order does not change); the others go to the shard tail."""
        emb = self._as_rows(embeddings)
        n = emb.shape[0]
        if docs is not None and self.keep_payload and len(docs) != n:
            # zip() in main.py:318 silently truncates to the shorter of the two
            n = min(n, len(docs))
            emb = emb[:n]
        with self._lock:
            if not (self.keep_payload and docs is not None):
                self._grow_locked(self._rows + n)
                self.add_device_rows(emb, _locked=True)
                return
            # main.py:325: i enumerates the rows of THIS call
            new_ids = list(ids[:n]) if ids is not None else [f"{docs[i]['doc_id']}_{i}" for i in range(n)]
            last = {}                                    # the same _id twice in one call: the last one wins
            for i, _id in enumerate(new_ids):
                last[_id] = i
            fresh = [i for i, _id in enumerate(new_ids) if last[_id] == i and _id not in self._row_of_id]
            over = [i for i, _id in enumerate(new_ids) if last[_id] == i and _id in self._row_of_id]
            if over:
                rows = torch.tensor([self._row_of_id[new_ids[i]] for i in over], dtype=torch.int64, device=self.device)
                with torch.cuda.device(self.device):
                    src = torch.from_numpy(np.ascontiguousarray(emb[over])).to(self.device)
                    fresh_rows = ops.normalize_cast(src, self.dtype)
                    self._shard.index_copy_(0, rows, fresh_rows)
                    if self.prefilter:
                        t8, tm = ops.quantize_rows(fresh_rows)
                        self._coarse8.index_copy_(0, rows, t8)
                        self._coarse_meta.index_copy_(0, rows, tm)
                    torch.cuda.current_stream(self.device).synchronize()
                for i in over:
                    self._docs[self._row_of_id[new_ids[i]]] = {"doc_id": docs[i]["doc_id"], "text": docs[i]["text"]}
            if fresh:
                base = self._rows
                while len(self._docs) < base:            # rows added without payload (add_device_rows)
                    self._docs.append({"doc_id": str(len(self._docs)), "text": ""})
                    self._ids.append(str(len(self._ids)))
                self._grow_locked(base + len(fresh))
                # rows first into the shard tail (not yet visible), then the payload tables, and the
                # row count LAST: a concurrent search never returns a row whose payload is missing
                self.add_device_rows(emb if len(fresh) == n else np.ascontiguousarray(emb[fresh]),
                                     _locked=True, _publish=False)
                for j, i in enumerate(fresh):
                    d = docs[i]
                    self._docs.append({"doc_id": d["doc_id"], "text": d["text"]})
                    self._ids.append(new_ids[i])
                    self._row_of_id[new_ids[i]] = base + j
                self._publish_rows(base + len(fresh))

    def _publish_rows(self, rows: int) -> None:
        self._rows = rows
        self._graphs.clear()                             # captured row count / shard pointer are stale

    def add_device_rows(self, emb, _locked: bool = False, _publish: bool = True) -> None:
        """Append rows without payload.  `emb`: host ndarray / CPU tensor / CUDA fp32 tensor
        [n,1024], un-normalised.  K1 writes straight into the shard tail, then the new row
        count is published."""
        if not _locked:
            with self._lock:
                self._grow_locked(self._rows + int(emb.shape[0]))
                return self.add_device_rows(emb, _locked=True, _publish=_publish)
        if isinstance(emb, np.ndarray):
            emb = torch.from_numpy(emb)
        n = int(emb.shape[0])
        if n == 0:
            return
        base = self._rows
        with torch.cuda.device(self.device):
            if emb.is_cuda:
                step = 1 << 18
                for lo in range(0, n, step):
                    hi = min(n, lo + step)
                    ops.normalize_cast(emb[lo:hi].contiguous(), self.dtype, out=self._shard[base + lo: base + hi])
            else:
                self._ingest_host_rows(emb, base, n)
            if self.prefilter:                           # the coarse copy of the new rows (K1q)
                ops.quantize_rows(self._shard[base: base + n], out=(self._coarse8, self._coarse_meta), row0=base)
            torch.cuda.current_stream(self.device).synchronize()
        if _publish:
            self._publish_rows(base + n)

    def _ingest_host_rows(self, emb: torch.Tensor, base: int, n: int, chunk: int = 1 << 15) -> None:
        """Host rows -> shard: two fp32 staging blocks on the device (128 MB each); the copy of
        block i+1 runs on a copy stream while K1 normalises block i into the shard tail, so a
        pinned source is ingested at PCIe speed (a pageable one at the host's memcpy speed)."""
        dev = self.device
        emb = emb.contiguous()
        if emb.dtype != torch.float32:
            emb = emb.float()
        compute = torch.cuda.current_stream(dev)
        copy = torch.cuda.Stream(device=dev)
        m = min(chunk, n)
        stage = [torch.empty((m, EMBED_DIM), dtype=torch.float32, device=dev) for _ in range(2 if n > m else 1)]
        ev_cp = [torch.cuda.Event() for _ in stage]
        ev_k1 = [torch.cuda.Event() for _ in stage]
        copy.wait_stream(compute)
        for i, lo in enumerate(range(0, n, m)):
            hi = min(n, lo + m)
            s = i % len(stage)
            with torch.cuda.stream(copy):
                if i >= len(stage):
                    copy.wait_event(ev_k1[s])            # K1 has consumed this block's previous content
                stage[s][: hi - lo].copy_(emb[lo:hi], non_blocking=True)
                ev_cp[s].record(copy)
            compute.wait_event(ev_cp[s])
            ops.normalize_cast(stage[s][: hi - lo], self.dtype, out=self._shard[base + lo: base + hi])
            ev_k1[s].record(compute)
        for t in stage:                                  # freed on `compute`, last used there too
            t.record_stream(copy)

    def clear(self) -> None:
        """Drop every row (capacity is kept): the index is empty again, `has_any_data()` False."""
        with self._lock:
            self._rows = 0
            self._docs.clear()
            self._ids.clear()
            self._row_of_id.clear()
            self._graphs.clear()

    # ------------------------------------------------------------------- search
    @staticmethod
    def _as_rows(x) -> np.ndarray:
        a = np.asarray(x)
        if a.ndim == 1:
            a = a[None, :]
        if a.ndim != 2 or a.shape[1] != EMBED_DIM:
            raise ValueError(f"expected [n,{EMBED_DIM}] embeddings, got {a.shape}")
        return np.ascontiguousarray(a, dtype=np.float32)

    def _stage_queries(self, q: np.ndarray) -> torch.Tensor:
        t = torch.from_numpy(q)
        if t.is_pinned():                                # caller already holds page-locked memory
            return t.to(self.device, non_blocking=True)
        b = q.shape[0]
        if self._pinned_q is None or self._pinned_q.shape[0] < b:
            self._pinned_q = torch.empty((max(b, 64), EMBED_DIM), dtype=torch.float32).pin_memory()
        self._pinned_q[:b].copy_(torch.from_numpy(q))
        return self._pinned_q[:b].to(self.device, non_blocking=True)

    def search_batch(self, query_emb: np.ndarray, k: int = 3) -> Tuple[np.ndarray, np.ndarray]:
        """Batched form of `search`: host fp32 [B,1024] in, host (scores [B,k] fp32,
        rows [B,k] int64) out, best-first; empty slots are (-inf, -1)."""
        q = self._as_rows(query_emb)
        with self._search_lock, torch.cuda.device(self.device):
            qd = self._stage_queries(q)
            buf, s, i = ops.packed_topk_out(self.device, q.shape[0], k)
            self.search_device(qd, k, out=(s, i))
            return self._fetch_packed(buf, q.shape[0], k)

    def _fetch_packed(self, buf: torch.Tensor, b: int, k: int) -> Tuple[np.ndarray, np.ndarray]:
        """ONE device->host copy of a packed (rows, scores) buffer into reusable pinned memory."""
        nbytes = buf.numel()
        if self._pinned_out is None or self._pinned_out.numel() < nbytes:
            self._pinned_out = torch.empty((max(nbytes, 4096),), dtype=torch.uint8).pin_memory()
        host = self._pinned_out[:nbytes]
        host.copy_(buf, non_blocking=True)
        torch.cuda.current_stream(self.device).synchronize()
        arr = host.numpy()
        rows = arr[: b * k * 8].view(np.int64).reshape(b, k).copy()
        scores = arr[b * k * 8:].view(np.float32).reshape(b, k).copy()
        return scores, rows

    def search_batches(self, batches, k: int = 3, depth: int = 2, device_fn=None):
        """Streaming form of `search_batch` for a sequence of query batches (host fp32 [B,1024]
        each): a generator that yields one `(scores [B,k], rows [B,k])` per input batch, in
        order, with identical contents.  The host->device copy of batch j+1 and the
        device->host copy of batch j-1 run on copy streams while batch j is being scored, so the
        GPU never waits for PCIe or for Python (ops.stream_pipeline).  `device_fn(q_dev, k, out)`
        replaces the local scan (ShardedCorpusIndex passes scan + exchange)."""
        fn = device_fn if device_fn is not None else (lambda qd, kk, out: self.search_device(qd, kk, out=out))

        def launch(qd: torch.Tensor, buf: torch.Tensor) -> None:
            b = qd.shape[0]
            fn(qd, k, (buf[b * k * 8:].view(torch.float32).view(b, k),
                       buf[: b * k * 8].view(torch.int64).view(b, k)))

        def unpack(arr: np.ndarray, b: int):
            rows = arr[: b * k * 8].view(np.int64).reshape(b, k).copy()
            scores = arr[b * k * 8:].view(np.float32).reshape(b, k).copy()
            return scores, rows

        return ops.stream_pipeline(self.device, batches, self._as_rows, lambda b: b * k * 12,
                                   launch, unpack, depth)

    @staticmethod
    def can_fuse_exchange(b: int) -> bool:
        """Whether `search_device` answers a batch of `b` queries with a scan whose last CTA can
        carry the sharded mode's exchange (`xchg=`): one or two queries (the fused exact scan, or
        the prefiltered scan when the index keeps its int8 copy).  Depends on nothing but `b`, so
        every rank of a sharded index takes the same path."""
        return 1 <= b <= 2

    def search_device(self, q_dev: torch.Tensor, k: int, idx_offset: int = 0, out=None, xchg=None,
                      queries_ready: bool = False) -> Tuple[torch.Tensor, torch.Tensor]:
        """Device-resident form: fp32 CUDA queries [B,1024] (un-normalised) -> CUDA
        (scores, rows).  No synchronisation.  `xchg` (ShardedCorpusIndex): fuse the exchange with
        the other ranks into the scan -- only when `can_fuse_exchange(B)`.  `queries_ready=True`:
        `q_dev` was complete before the previous kernel on this stream was launched (a resident
        query batch); one- and two-query scans may then overlap the previous scan's tail."""
        rows = self._rows
        shard = self._shard if self._shard is not None else self.shard
        q_dev = q_dev.contiguous()
        c8, cm = self._coarse8, self._coarse_meta
        if self.prefilter and c8 is not None and 1 <= q_dev.shape[0] <= 2 and 0 < rows <= c8.shape[0] \
                and q_dev.dtype == torch.float32:
            # K3p: int8 prefilter + exact rescoring -- the exact scan's results at half its bytes;
            # both passes normalise the raw query themselves (no K1 launch)
            return ops.search_gemv_prefiltered(shard, c8, cm, q_dev, k, idx_offset=idx_offset, n=rows, out=out,
                                               xchg=xchg, queries_ready=queries_ready)
        if self.prefilter and c8 is not None and q_dev.shape[0] > 2 and 0 < rows <= c8.shape[0] \
                and q_dev.dtype == torch.float32 and xchg is None \
                and k <= (nat.SQE_MAX_K_BATCHED if self.dtype == "fp32" else self.K2P_MAX_K) \
                and ops.k2p_pays(rows, q_dev.shape[0], self.dtype):
            # K2p: the batch form of the same idea on the int8 tensor cores.  16-bit shards: up to
            # K2P_MAX_K (beyond that the exact pass outweighs the cheaper scan and K2 is as fast) and
            # from the shard size where it pays (ops.k2p_pays); fp32 shards: always (their only
            # other batch path is one streaming pass per query)
            return ops.search_batched_prefiltered(shard, c8, cm, q_dev, k, idx_offset=idx_offset, n=rows, out=out)
        if q_dev.dtype == torch.float32 and (q_dev.shape[0] == 1 or (xchg is not None and q_dev.shape[0] == 2)):
            # the reference's own case (one query, main.py:355): normalise + scan in ONE launch
            return ops.search_gemv(shard, q_dev, k, idx_offset=idx_offset, n=rows, out=out, xchg=xchg,
                                   queries_ready=queries_ready)
        if xchg is not None:
            raise ValueError("this batch cannot carry a fused exchange (see can_fuse_exchange)")
        qn = ops.normalize_cast(q_dev, self.dtype)                       # main.py:353-354
        return ops.topk(shard, qn, k, idx_offset=idx_offset, n=rows, out=out)

    def search(self, query_emb: np.ndarray, k: int = 3) -> List[Tuple[Dict[str, str], float]]:
        """main.py:347-373."""
        if query_emb is None or getattr(query_emb, "size", 0) == 0:      # main.py:350-351
            return []
        try:
            q = self._as_rows(query_emb)[:1]                             # main.py:355 sends row 0
            scores, rows = self._search_one(q, k)
            return self.hits_from_rows(scores, rows)
        except Exception as e:                                           # main.py:371-373
            if self.strict:
                raise
            print(f"[GpuCorpusIndex] Search error: {e}")
            return []

    def _search_one(self, q: np.ndarray, k: int) -> Tuple[np.ndarray, np.ndarray]:
        """The reference's call shape (one query): replay a captured CUDA graph (H2D, fused
        normalise + scan + top-k, D2H) when possible, else the eager path.

        Capture policy: a graph is captured at the SECOND search of an unchanged index state (the
        first one runs eagerly -- while rows are still arriving every search would otherwise pay a
        warm-up scan + a capture), under the ingest lock (no rows are added, no shard is regrown
        while the capture is open) and in thread-local capture mode.  A failed capture is retried
        after `_GRAPH_RETRY_S`, not never."""
        if self.use_graphs and self._rows > 0 and 1 <= k <= nat.SQE_MAX_K_GEMV:
            with self._search_lock:
                g = self._graphs.get(k)
                state = (self._rows, self._shard.data_ptr(), self.prefilter)
                if g is not None and (g.rows, g.shard_ptr) != state[:2]:
                    g = None
                    self._graphs.pop(k, None)
                if g is None and self._graph_seen.get(k) != state:
                    self._graph_seen[k] = state          # first search of this state: eager
                elif g is None and time.monotonic() >= self._graph_retry_at and self._lock.acquire(blocking=False):
                    try:                                  # ingest in progress -> lock busy -> stay eager
                        if (self._rows, self._shard.data_ptr(), self.prefilter) == state:
                            g = ops.SingleQueryGraph(self._shard, self._rows, k, coarse=(
                                (self._coarse8, self._coarse_meta, self.dtype) if self.prefilter else None))
                            self._graphs[k] = g
                    except Exception as e:               # capture not possible right now: eager, retry later
                        print(f"[GpuCorpusIndex] CUDA graph capture failed ({e}); eager launches for "
                              f"{self._GRAPH_RETRY_S:.0f} s")
                        self._graph_retry_at = time.monotonic() + self._GRAPH_RETRY_S
                        g = None
                    finally:
                        self._lock.release()
                if g is not None:
                    return g.run(q[0])
        scores, rows = self.search_batch(q, k)
        return scores[0], rows[0]

    def hits_from_rows(self, scores, rows) -> List[Tuple[Dict[str, str], float]]:
        """(score, row) arrays of ONE query -> the reference's hit list (main.py:364-367)."""
        if not isinstance(scores, list):                 # numpy in, Python floats / ints out
            scores, rows = np.asarray(scores).tolist(), np.asarray(rows).tolist()
        if self.return_embedding:
            src = self._source
        elif self.keep_payload:                          # a fresh dict per hit, like hit["_source"]
            docs, n_docs = self._docs, len(self._docs)
            src = lambda r: dict(docs[r]) if r < n_docs else {"doc_id": str(r), "text": ""}
        else:
            src = lambda r: {"doc_id": str(r), "text": ""}
        if self.score_mode == "opensearch":
            return [(src(r), 1.0 / (2.0 - s)) for s, r in zip(scores, rows) if r >= 0]
        return [(src(r), s) for s, r in zip(scores, rows) if r >= 0]

    def _source(self, row: int) -> Dict[str, str]:
        if self.keep_payload and row < len(self._docs):
            src = dict(self._docs[row])
        else:
            src = {"doc_id": str(row), "text": ""}
        if self.return_embedding:                                        # main.py:326-330
            stored = self._shard[row].float().cpu()
            if self.dtype == "bf16x2":
                stored = stored[:EMBED_DIM] + stored[EMBED_DIM:]          # hi + lo
            src["embedding"] = stored.tolist()
        return src

    def doc_id_of(self, row: int) -> str:
        return self._ids[row]

    # ------------------------------------------- the reference's OpenSearch documents
    def export_bulk_actions(self, chunk_rows: int = 4096):
        """Yield one OpenSearch bulk action per stored row in the reference's own format
        (main.py:318-331): `{"_op_type": "index", "_index", "_id", "_source": {"doc_id", "text",
        "embedding": [1024 floats]}}` -- what `helpers.bulk` needs to rebuild the reference's
        k-NN index from this shard.  The embedding is the STORED (normalised, rounded) row."""
        rows = self._rows
        for lo in range(0, rows, chunk_rows):
            blk = self._shard[lo: min(rows, lo + chunk_rows)].float().cpu()
            if self.dtype == "bf16x2":
                blk = blk[:, :EMBED_DIM] + blk[:, EMBED_DIM:]
            blk = blk.numpy()
            for j in range(blk.shape[0]):
                r = lo + j
                have = self.keep_payload and r < len(self._docs)
                yield {"_op_type": "index", "_index": self.index_name,
                       "_id": self._ids[r] if have else str(r),
                       "_source": {"doc_id": self._docs[r]["doc_id"] if have else str(r),
                                   "text": self._docs[r]["text"] if have else "",
                                   "embedding": blk[j].tolist()}}

    def import_bulk_actions(self, actions, batch_rows: int = 4096) -> int:
        """Ingest documents in the reference's OpenSearch format -- bulk actions (main.py:318-331)
        or search/scroll hits (`{"_id", "_source": {...}}`) -- e.g. to move an index the reference
        built into HBM.  `_id` is kept as it is (a known `_id` overwrites its row, as in OpenSearch).
        Embeddings go through K1 like any other ingest (re-normalising a unit row changes it by at
        most one rounding).  Returns the number of documents processed."""
        total = 0
        emb, docs, ids = [], [], []

        def flush():
            nonlocal total
            if not emb:
                return
            self._add(np.asarray(emb, dtype=np.float32), docs, id_from_global_row=True, ids=list(ids))
            total += len(emb)
            emb.clear(); docs.clear(); ids.clear()

        for a in actions:
            src = a["_source"]
            emb.append(src["embedding"])
            docs.append({"doc_id": src["doc_id"], "text": src["text"]})
            ids.append(a.get("_id", f"{src['doc_id']}_{len(ids)}"))
            if len(emb) >= batch_rows:
                flush()
        flush()
        return total

    # -------------------------------------------------------------- persistence
    # The reference's only "resume" is skipping ingest when the index already has data
    # (main.py:300-307, :422-424); its persistence is whatever OpenSearch keeps.  Here the packed
    # shard is written as it sits in HBM (stored rows, no re-normalisation on load) next to the
    # payload table, so a restarted process answers `has_any_data()` the same way.
    FORMAT_VERSION = 1

    def save(self, path: str, chunk_rows: int = 1 << 17) -> None:
        """Write `path`/shard.bin (raw stored rows), meta.json, payload.json."""
        os.makedirs(path, exist_ok=True)
        with self._lock:
            rows = self._rows
            row_bytes = ops.ROW_BYTES[self.dtype]
            with open(os.path.join(path, "shard.bin"), "wb") as f:
                for lo in range(0, rows, chunk_rows):
                    blk = self._shard[lo: min(rows, lo + chunk_rows)]
                    raw = blk.contiguous().view(torch.uint8).cpu().numpy()
                    f.write(raw.tobytes())
            with open(os.path.join(path, "payload.json"), "w") as f:
                json.dump({"docs": self._docs, "ids": self._ids}, f)
            with open(os.path.join(path, "meta.json"), "w") as f:
                json.dump({"format": self.FORMAT_VERSION, "rows": rows, "dim": EMBED_DIM,
                           "dtype": self.dtype, "bytes": rows * row_bytes,
                           "index_name": self.index_name, "normalised": True}, f)

    @classmethod
    def load(cls, path: str, *, device: Optional[torch.device] = None, chunk_rows: int = 1 << 17,
             **kwargs) -> "GpuCorpusIndex":
        """Rebuild an index from `save()` output: the stored rows go to HBM bit for bit."""
        with open(os.path.join(path, "meta.json")) as f:
            meta = json.load(f)
        if meta.get("format") != cls.FORMAT_VERSION or meta.get("dim") != EMBED_DIM:
            raise ValueError(f"unsupported shard file {path}: {meta}")
        index = cls(None, meta.get("index_name", ""), dtype=meta["dtype"], device=device, **kwargs)
        rows = int(meta["rows"])
        row_bytes = ops.ROW_BYTES[index.dtype]
        size = os.path.getsize(os.path.join(path, "shard.bin"))
        if size != rows * row_bytes:
            raise ValueError(f"shard.bin has {size} bytes, expected {rows * row_bytes}")
        with index._lock:
            index._grow_locked(max(rows, 1))
            with open(os.path.join(path, "shard.bin"), "rb") as f:
                for lo in range(0, rows, chunk_rows):
                    n = min(chunk_rows, rows - lo)
                    raw = np.frombuffer(f.read(n * row_bytes), dtype=np.uint8)
                    dst = index._shard[lo: lo + n].view(torch.uint8)
                    dst.copy_(torch.from_numpy(raw.copy()).view(n, row_bytes))
            if index.prefilter and rows:
                ops.quantize_rows(index._shard[:rows], out=(index._coarse8, index._coarse_meta))
            torch.cuda.current_stream(index.device).synchronize()
            index._rows = rows
            if index.keep_payload and os.path.isfile(os.path.join(path, "payload.json")):
                with open(os.path.join(path, "payload.json")) as f:
                    payload = json.load(f)
                index._docs = payload["docs"]
                index._ids = payload["ids"]
                index._row_of_id = {_id: r for r, _id in enumerate(index._ids)}
        return index
