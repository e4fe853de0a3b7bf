"""GPU-resident mirror of the reference's Redis LFU query cache
(/root/reference/app/main.py:56-128): drop-in for `lfu_cache_get` / `lfu_cache_put`.

The reference keeps a Redis list of JSON entries `{"embedding", "response", "freq"}`
(main.py:123) with the newest entry at index 0 (LPUSH, main.py:128); every lookup
re-parses every entry (385 ms at 1000 entries).  Here the embeddings live in HBM as unit
rows in *list order* -- row `head + i` is list index `i` -- so "first maximum wins"
(strict `>`, main.py:84) is the kernel's "lowest row wins" rule, and a lookup is one K1 +
one K5 launch.  `response` / `freq` stay in a host list in the same order.

LFU eviction (main.py:101-118) removes the FIRST entry with the minimal freq; the rows
in front of it slide down by one so list order is preserved.  Optionally every mutation is
written through to a Redis client in the reference's own entry format, so a reference
process can keep reading the same list.
"""
from __future__ import annotations

import json
import math
import threading
from typing import List, Optional, Tuple

import numpy as np
import torch

from . import _native as nat
from . import ops

REDIS_MAX_ITEMS = 1000          # main.py:42
REDIS_CACHE_LIST = "query_cache_lfu"   # main.py:43
CACHE_SIM_THRESHOLD = 0.96      # main.py:44


def cosine_similarity(a, b, *, device: Optional[torch.device] = None) -> float:
    """Drop-in for the reference's `cosine_similarity(a, b)` (main.py:59-64) on the GPU: `b` becomes
    a one-row fp32 unit shard (K1) and `a` is scored against it by the fused normalise + scan
    kernel (`sqe_search_gemv`, k = 1).  A zero-norm argument gives 0.0 like main.py:62-63 (a zero
    row stays zero under x/(||x||+1e-9)).  Agrees with the numpy expression within the fp32
    tolerance (1e-5; measured ~1e-7).  One pair per call is launch-bound -- the handlers never
    call it once `lfu_cache_get` is served by `GpuQueryCache`; it exists so that every name of the
    path has a GPU counterpart."""
    dev = torch.device(device) if device is not None else torch.device("cuda", 0)
    va, vb = GpuQueryCache._row0(a), GpuQueryCache._row0(b)
    if va is None or vb is None:
        raise ValueError(f"expected two [{nat.SQE_DIM}] embeddings")
    na, nb = float(np.linalg.norm(va)), float(np.linalg.norm(vb))
    if na == 0.0 or nb == 0.0:                           # main.py:62-63
        return 0.0
    # The reference divides by the norms themselves, the kernels by (norm + 1e-9): bring both
    # norms into [0.5, 1) with an exact power-of-two scaling so that the 1e-9 is immaterial for
    # tiny (or huge) vectors too; the cosine is scale-invariant.
    if np.isfinite(na) and np.isfinite(nb):
        va = np.ldexp(va, -math.frexp(na)[1]).astype(np.float32)
        vb = np.ldexp(vb, -math.frexp(nb)[1]).astype(np.float32)
    with torch.cuda.device(dev):
        row = ops.normalize_cast(torch.from_numpy(vb[None, :]).to(dev), "fp32")
        scores, _ = ops.search_gemv(row, torch.from_numpy(va[None, :]).to(dev), 1)
        return float(scores.cpu()[0, 0])


class GpuQueryCache:
    def __init__(self, max_items: int = REDIS_MAX_ITEMS, threshold: float = CACHE_SIM_THRESHOLD,
                 *, dtype: str = "fp32", device: Optional[torch.device] = None,
                 redis_client=None, list_name: str = REDIS_CACHE_LIST, keep_raw: bool = None,
                 use_graphs: bool = True):
        if dtype not in ops.TORCH_DTYPES:
            raise ValueError(f"dtype must be one of {sorted(ops.TORCH_DTYPES)}")
        self.max_items = int(max_items)
        self.threshold = float(threshold)
        self.dtype = dtype
        self.device = torch.device(device) if device is not None else torch.device("cuda", 0)
        self.redis = redis_client
        self.list_name = list_name
        # raw embeddings are only needed to write the reference's JSON entry
        self.keep_raw = (redis_client is not None) if keep_raw is None else keep_raw
        self._lock = threading.RLock()          # also guards the pinned staging buffers
        self._buf = torch.zeros((self.max_items, ops.ROW_ELEMS[dtype]), dtype=ops.TORCH_DTYPES[dtype],
                                device=self.device)
        self._head = self.max_items             # live rows are [_head, max_items)
        self._entries: List[dict] = []          # list order, index 0 = newest
        self.use_graphs = use_graphs
        self._graph = None                       # captured single-query lookup; dropped on every mutation
        self._graph_key = None                   # cache state seen by the last eager lookup
        self._pinned = torch.empty((1, nat.SQE_DIM), dtype=torch.float32).pin_memory()
        self._pinned_out = torch.empty((4096,), dtype=torch.uint8).pin_memory()
        self._pinned_qb: Optional[torch.Tensor] = None

    def __len__(self) -> int:
        return len(self._entries)

    # ---------------------------------------------------------------- helpers
    @staticmethod
    def _row0(query_emb) -> Optional[np.ndarray]:
        if query_emb is None or getattr(query_emb, "size", 0) == 0:
            return None
        a = np.asarray(query_emb, dtype=np.float32)
        if a.ndim == 2:
            a = a[0]                             # main.py:73 uses row 0
        if a.shape != (nat.SQE_DIM,):
            raise ValueError(f"expected [1,{nat.SQE_DIM}] query embedding, got {np.shape(query_emb)}")
        return np.ascontiguousarray(a)

    def _to_device(self, vec: np.ndarray) -> torch.Tensor:
        self._pinned[0].copy_(torch.from_numpy(vec))
        return self._pinned.to(self.device, non_blocking=True)

    def _entry_json(self, e: dict) -> str:
        raw = e.get("raw")
        if raw is None:
            # entry without its raw embedding (bulk_load): write the stored unit row -- the same
            # cosine for every reader, the reference normalises at lookup (main.py:59-64)
            i = next(j for j, x in enumerate(self._entries) if x is e)
            row = self._buf[self._head + i].float().cpu()
            if self.dtype == "bf16x2":
                row = row[: nat.SQE_DIM] + row[nat.SQE_DIM:]
            raw = e["raw"] = row.tolist()
        return json.dumps({"embedding": raw, "response": e["response"], "freq": e["freq"]})

    # ------------------------------------------------------------------- get
    def lookup(self, query_emb) -> Tuple[int, float, bool]:
        """(list index, similarity, hit) of the best entry; (-1, -1.0, False) when empty."""
        vec = self._row0(query_emb)
        if vec is None or not self._entries:
            return -1, -1.0, False
        with self._lock, torch.cuda.device(self.device):
            row, sim = self._top1(vec)
        best_sim, best_index = -1.0, -1                  # main.py:74-75
        if row >= 0 and sim > best_sim:                  # main.py:84 (strict '>')
            best_sim, best_index = sim, row
        hit = best_index >= 0 and not (best_sim < self.threshold)    # main.py:89
        return best_index, best_sim, hit

    def _top1(self, vec: np.ndarray) -> Tuple[int, float]:
        """(row, similarity) of the best live entry for one raw query.  ONE kernel: normalise the
        query + scan + top-1 (sqe_search_gemv, k = 1) -- replayed from a captured CUDA graph
        (H2D, kernel, D2H) when possible."""
        live = self._buf[self._head:]
        if self.use_graphs:
            g = self._graph
            stale = g is None or g.rows != len(self._entries) or g.shard_ptr != live.data_ptr()
            if stale and self._graph_key != (len(self._entries), live.data_ptr()):
                # first lookup of this cache state: launch eagerly.  The reference's handler
                # alternates get (miss) -> put (main.py:493, :547); capturing a graph per state
                # would cost more than it saves.  A second lookup of the same state captures.
                self._graph_key = (len(self._entries), live.data_ptr())
                g = self._graph = None
            elif stale:
                try:
                    g = self._graph = ops.SingleQueryGraph(live, len(self._entries), 1)
                except Exception as e:                       # capture not possible here: stay eager
                    print(f"[GpuQueryCache] CUDA graph capture failed ({e}); using eager launches")
                    self.use_graphs = False
                    g = None
            if g is not None:
                s_, r_ = g.run(vec)
                return int(r_[0]), float(s_[0])
        buf, s, i = ops.packed_topk_out(self.device, 1, 1)
        ops.search_gemv(live, self._to_device(vec), 1, n=len(self._entries), out=(s, i))
        self._pinned_out[:12].copy_(buf, non_blocking=True)     # one 12-byte copy back
        torch.cuda.current_stream(self.device).synchronize()
        raw = self._pinned_out[:12].numpy()
        return int(raw[:8].view(np.int64)[0]), float(raw[8:12].view(np.float32)[0])

    def get(self, query_emb) -> Optional[str]:
        """lfu_cache_get, main.py:67-98."""
        with self._lock:
            if not self._entries:                                    # main.py:70-71
                return None
            idx, _sim, hit = self.lookup(query_emb)
            if not hit:                                              # main.py:89-90
                return None
            e = self._entries[idx]
            e["freq"] = e.get("freq", 1) + 1                         # main.py:94
            if self.redis is not None:
                self.redis.lset(self.list_name, idx, self._entry_json(e))   # main.py:95
            return e["response"]

    def lookup_batch(self, queries: np.ndarray, path: int = 0
                     ) -> Tuple[np.ndarray, np.ndarray, np.ndarray]:
        """Streaming form (BASELINE config 5): host fp32 [B,1024] -> host
        (idx int32 [B], score fp32 [B], hit uint8 [B]).  Does not touch `freq`."""
        q = np.ascontiguousarray(np.asarray(queries, dtype=np.float32))
        if q.ndim != 2 or q.shape[1] != nat.SQE_DIM:
            raise ValueError("expected [B,1024] queries")
        b = q.shape[0]
        with self._lock, torch.cuda.device(self.device):
            if self._pinned_qb is None or self._pinned_qb.shape[0] < b:
                self._pinned_qb = torch.empty((max(b, 64), nat.SQE_DIM), dtype=torch.float32).pin_memory()
            self._pinned_qb[:b].copy_(torch.from_numpy(q))
            qd = self._pinned_qb[:b].to(self.device, non_blocking=True)
            buf, idx, score, hit = ops.packed_cache_out(self.device, b)
            self.lookup_device(qd, path, out=(idx, score, hit))
            if self._pinned_out.numel() < b * 9:
                self._pinned_out = torch.empty((b * 9,), dtype=torch.uint8).pin_memory()
            host = self._pinned_out[: b * 9]
            host.copy_(buf, non_blocking=True)             # ONE device->host copy
            torch.cuda.current_stream(self.device).synchronize()
            raw = host.numpy()
            return (raw[: b * 4].view(np.int32).copy(), raw[b * 4: b * 8].view(np.float32).copy(),
                    raw[b * 8:].copy())

    def lookup_batches(self, batches, path: int = 0, depth: int = 2):
        """`lookup_batch` for a stream of query batches (BASELINE config 5 is exactly this):
        a generator yielding `(idx, score, hit)` per batch, in order, identical to `lookup_batch`;
        the copies of neighbouring batches overlap the scan (ops.stream_pipeline).  The cache must
        not be mutated while the generator is being consumed."""
        def as_rows(x) -> np.ndarray:
            q = np.ascontiguousarray(np.asarray(x, dtype=np.float32))
            if q.ndim != 2 or q.shape[1] != nat.SQE_DIM:
                raise ValueError("expected [B,1024] queries")
            return q

        def launch(qd: torch.Tensor, buf: torch.Tensor) -> None:
            b = qd.shape[0]
            self.lookup_device(qd, path, out=(buf[: b * 4].view(torch.int32),
                                              buf[b * 4: b * 8].view(torch.float32), buf[b * 8:]))

        def unpack(raw: np.ndarray, b: int):
            return (raw[: b * 4].view(np.int32).copy(), raw[b * 4: b * 8].view(np.float32).copy(),
                    raw[b * 8:].copy())

        return ops.stream_pipeline(self.device, batches, as_rows, lambda b: b * 9, launch, unpack, depth)

    def lookup_device(self, q_dev: torch.Tensor, path: int = 0, out=None):
        qn = ops.normalize_cast(q_dev.contiguous(), self.dtype)
        return ops.cache_top1(self._buf[self._head:], qn, self.threshold, path=path,
                              n=len(self._entries), out=out)

    # ------------------------------------------------------------------- put
    def _remove_least_frequent_item(self) -> None:
        """main.py:101-118: first entry with the minimal freq (strict '<')."""
        if not self._entries:
            return
        min_freq = float("inf")
        min_index = -1
        for i, e in enumerate(self._entries):
            f = e.get("freq", 1)
            if f < min_freq:
                min_freq = f
                min_index = i
        if min_index < 0:
            return
        e = self._entries[min_index]
        if self.redis is not None:
            self.redis.lrem(self.list_name, 1, self._entry_json(e))  # main.py:117
        self._entries.pop(min_index)
        h = self._head
        if min_index > 0:
            # rows [h, h+min_index) slide to [h+1, h+min_index+1): list order is kept
            self._buf[h + 1: h + min_index + 1] = self._buf[h: h + min_index].clone()
        self._head = h + 1

    def put(self, query_emb, response: str) -> None:
        """lfu_cache_put, main.py:121-128.  The reference stores the RAW embedding and
        normalises inside cosine_similarity at every lookup; storing the unit row once is the
        same cosine."""
        vec = self._row0(query_emb)
        if vec is None:
            return
        with self._lock:
            if len(self._entries) >= self.max_items:                 # main.py:125-126
                self._remove_least_frequent_item()
            if self._head == 0:
                raise RuntimeError("cache buffer full (max_items reached with nothing evictable)")
            with torch.cuda.device(self.device):
                self._head -= 1
                ops.normalize_cast(self._to_device(vec), self.dtype,
                                   out=self._buf[self._head: self._head + 1])
                torch.cuda.current_stream(self.device).synchronize()   # pinned staging reused
            e = {"response": response, "freq": 1}
            if self.keep_raw:
                e["raw"] = vec.tolist()
            self._entries.insert(0, e)
            if self.redis is not None:
                self.redis.lpush(self.list_name, self._entry_json(e))  # main.py:128

    def bulk_load(self, embeddings, responses=None) -> None:
        """Fill an empty cache with many entries at once (config-5 sized caches).
        Row i becomes list index i.  `embeddings`: host ndarray or CUDA fp32 tensor."""
        n = int(embeddings.shape[0])
        if self._entries:
            raise RuntimeError("bulk_load needs an empty cache")
        if n > self.max_items:
            raise ValueError("more entries than max_items")
        with self._lock, torch.cuda.device(self.device):
            self._head = self.max_items - n
            step = 1 << 18
            for lo in range(0, n, step):
                hi = min(n, lo + step)
                blk = embeddings[lo:hi]
                if isinstance(blk, np.ndarray):
                    blk = torch.from_numpy(np.ascontiguousarray(blk, dtype=np.float32))
                blk = blk.to(self.device).contiguous()
                ops.normalize_cast(blk, self.dtype, out=self._buf[self._head + lo: self._head + hi])
            torch.cuda.current_stream(self.device).synchronize()
            self._entries = [{"response": (responses[i] if responses is not None else str(i)),
                              "freq": 1} for i in range(n)]

    def load_from_redis(self, redis_client=None, list_name: Optional[str] = None) -> int:
        """Warm start from the list a reference process has been writing: LRANGE the whole list
        (main.py:69), parse every entry the way lfu_cache_get does (main.py:79-82:
        `{"embedding": [1024 floats], "response": str, "freq": int}`), keep the list order (index 0
        = newest) and the `freq` counters.  The cache must be empty; nothing is written back.
        Returns the number of entries loaded."""
        client = redis_client if redis_client is not None else self.redis
        if client is None:
            raise ValueError("no Redis client")
        name = list_name if list_name is not None else self.list_name
        items = client.lrange(name, 0, -1)
        if not items:
            return 0
        if len(items) > self.max_items:
            raise ValueError(f"Redis list holds {len(items)} entries, max_items is {self.max_items}")
        parsed = [json.loads(it) for it in items]
        emb = np.asarray([e["embedding"] for e in parsed], dtype=np.float32)     # main.py:80
        if emb.ndim != 2 or emb.shape[1] != nat.SQE_DIM:
            raise ValueError(f"cache entries must hold {nat.SQE_DIM}-d embeddings, got {emb.shape}")
        self.bulk_load(emb, [e["response"] for e in parsed])
        with self._lock:
            for mine, theirs in zip(self._entries, parsed):
                mine["freq"] = theirs.get("freq", 1)                             # main.py:107
                if self.keep_raw:
                    mine["raw"] = theirs["embedding"]
        return len(parsed)

    # ------------------------------------------------------------- inspection
    def responses(self) -> List[str]:
        return [e["response"] for e in self._entries]

    def freqs(self) -> List[int]:
        return [e["freq"] for e in self._entries]
