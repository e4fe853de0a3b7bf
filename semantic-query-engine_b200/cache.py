"""GPU-resident mirror of the reference's Redis LFU query cache
(/root/reference/app/main.py:56-128): drop-in for `lfu_cache_get` / `lfu_cache_put`.

The reference keeps a Redis list of JSON entries `{"embedding", "response", "freq"}`
(main.py:123) with the newest entry at index 0 (LPUSH, main.py:128); every lookup
re-parses every entry (385 ms at 1000 entries).  Here the embeddings live in HBM as unit
rows in *list order* -- a lower row is a lower list index -- so "first maximum wins"
(strict `>`, main.py:84) is the kernels' "lowest row wins" rule, and a lookup is one fused
launch (b = 1) or K1 + K5 (batches).  `response` / `freq` stay on the host.

Mutation is O(1) amortised at any size (round 1 slid up to `max_items` rows per eviction):
  * the row buffer holds 2 x max_items rows; a new entry (list index 0) is written just BELOW
    the current head, so list order stays row order without moving anything;
  * LFU eviction (main.py:101-118: the FIRST entry with the minimal freq) takes the victim
    from a heap keyed (freq, newest first) in O(log n).  A victim within `_SLIDE_MAX` rows of
    the head -- the usual case, new entries have freq 1 -- is removed physically (the few rows
    in front of it slide down by one); a victim deep in the list becomes a tombstone: its row is
    zeroed (similarity exactly 0.0, below any positive threshold) and skipped on the host;
  * when the head reaches row 0 the live rows are packed to the top of the buffer in one
    device copy (once every >= max_items puts); tombstones disappear there.
Optionally every mutation is written through to a Redis client in the reference's own entry
format, so a reference process can keep reading the same list.

`prefilter=True` keeps the int8 copy of the rows (K1q) and answers BATCH lookups through K2p
(int8 tensor-core prefilter + exact rescoring): the best entry and its score are those of the
exact scan bit for bit, at half the bytes -- and it gives fp32-stored caches a tensor-core path.
"""
from __future__ import annotations

import bisect
import heapq
import json
import math
import threading
import time
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


class _Entry:
    __slots__ = ("response", "freq", "raw", "seq", "row")

    def __init__(self, response, freq, raw, seq, row):
        self.response, self.freq, self.raw, self.seq, self.row = response, freq, raw, seq, row


class GpuQueryCache:
    _SLIDE_MAX = 256            # evictions this close to the head slide rows instead of leaving a tombstone
    _GRAPH_RETRY_S = 5.0
    _STAGES = 8                 # pinned staging buffers of `put` (no stream sync per call)

    def __init__(self, max_items: int = REDIS_MAX_ITEMS, threshold: float = CACHE_SIM_THRESHOLD,
                 *, dtype: str = "fp32", device: Optional[torch.device] = None,
                 redis_client=None, list_name: str = REDIS_CACHE_LIST, keep_raw: bool = None,
                 use_graphs: bool = True, prefilter: bool = False):
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
        self.prefilter = bool(prefilter)
        self._lock = threading.RLock()          # also guards the pinned staging buffers
        self._cap = 2 * max(self.max_items, 1)  # rows of the buffer
        self._buf = torch.zeros((self._cap, ops.ROW_ELEMS[dtype]), dtype=ops.TORCH_DTYPES[dtype],
                                device=self.device)
        self._c8 = self._cm = None
        if self.prefilter:
            self._c8 = torch.zeros((self._cap, nat.SQE_DIM), dtype=torch.int8, device=self.device)
            self._cm = torch.zeros((self._cap, 4), dtype=torch.float32, device=self.device)
        self._head = self._cap                  # rows [_head, _cap) are the list, in list order
        self._slot: List[Optional[_Entry]] = [None] * self._cap     # row -> entry (None: free or tombstone)
        self._live = 0
        self._dead_rows: List[int] = []         # tombstones inside [_head, _cap), ascending
        self._heap: List[Tuple[int, int, int]] = []     # (freq, -seq, row), lazily invalidated
        self._seq = 0
        self.use_graphs = use_graphs
        self._graph = None                       # captured single-query lookup; dropped on every mutation
        self._graph_key = None                   # cache state seen by the last eager lookup
        self._graph_retry_at = 0.0
        self._version = 0                        # bumped by every mutation (graph staleness)
        self._stage = [torch.empty((1, nat.SQE_DIM), dtype=torch.float32).pin_memory() for _ in range(self._STAGES)]
        self._stage_ev = [None] * self._STAGES
        self._stage_i = 0
        self._pinned_out = torch.empty((4096,), dtype=torch.uint8).pin_memory()
        self._pinned_qb: Optional[torch.Tensor] = None

    def __len__(self) -> int:
        return self._live

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
        """One of a ring of pinned staging rows -> device, asynchronously (the ring slot is reused
        only after the copy that last used it has completed)."""
        i = self._stage_i
        self._stage_i = (i + 1) % self._STAGES
        ev = self._stage_ev[i]
        if ev is not None:
            ev.synchronize()
        self._stage[i][0].copy_(torch.from_numpy(vec))
        t = self._stage[i].to(self.device, non_blocking=True)
        ev = self._stage_ev[i] or torch.cuda.Event()
        ev.record(torch.cuda.current_stream(self.device))
        self._stage_ev[i] = ev
        return t

    def _scan(self) -> Tuple[torch.Tensor, int]:
        """The rows a lookup scans: (view starting at the head, number of rows incl. tombstones)."""
        return self._buf[self._head:], self._cap - self._head

    def scan_view(self) -> Tuple[torch.Tensor, int]:
        return self._scan()

    def _index_of_row(self, row: int) -> int:
        """List index (0 = newest) of the live entry stored at `row`."""
        return row - self._head - bisect.bisect_left(self._dead_rows, row)

    def _live_rows(self):
        slot = self._slot
        return (r for r in range(self._head, self._cap) if slot[r] is not None)

    def _entry_json(self, e: _Entry) -> str:
        raw = e.raw
        if raw is None:
            # entry without its raw embedding (bulk_load): write the stored unit row -- the same
            # cosine for every reader, the reference normalises at lookup (main.py:59-64)
            row = self._buf[e.row].float().cpu()
            if self.dtype == "bf16x2":
                row = row[: nat.SQE_DIM] + row[nat.SQE_DIM:]
            raw = e.raw = row.tolist()
        return json.dumps({"embedding": raw, "response": e.response, "freq": e.freq})

    def _first_live_best(self, vec: np.ndarray) -> Tuple[int, float]:
        """Rare: the scan's best row is a tombstone (similarity 0.0 beat every live entry).  Ask for
        as many rows as there are tombstones + 1 and take the best live one."""
        live, rows = self._scan()
        k = min(len(self._dead_rows) + 1, nat.SQE_MAX_K_GEMV, rows)
        s, i = ops.search_gemv(live, self._to_device(vec), k, n=rows)
        s, i = s.cpu().numpy()[0], i.cpu().numpy()[0]
        for sc, r in zip(s, i):
            if r >= 0 and self._slot[self._head + int(r)] is not None:
                return int(r), float(sc)
        # more tombstones than k: exact host-side fallback over the live rows only
        best_r, best_s = -1, -1.0
        q = ops.normalize_cast(self._to_device(vec), "fp32").cpu().numpy()[0]
        for r in self._live_rows():
            row = self._buf[r].float().cpu().numpy()
            if self.dtype == "bf16x2":
                row = row[: nat.SQE_DIM] + row[nat.SQE_DIM:]
            sc = float(np.dot(row, q))
            if sc > best_s:
                best_r, best_s = r - self._head, sc
        return best_r, best_s

    # ------------------------------------------------------------------- get
    def lookup(self, query_emb) -> Tuple[int, float, bool]:
        """(list index, similarity, hit) of the best entry; (-1, -1.0, False) when empty."""
        vec = self._row0(query_emb)
        if vec is None or not self._live:
            return -1, -1.0, False
        with self._lock, torch.cuda.device(self.device):
            rel, sim = self._top1(vec)
            if rel >= 0 and self._slot[self._head + rel] is None:      # a tombstone won (every live sim < 0)
                rel, sim = self._first_live_best(vec)
            row = self._head + rel if rel >= 0 else -1
        best_sim, best_index = -1.0, -1                  # main.py:74-75
        if row >= 0 and sim > best_sim:                  # main.py:84 (strict '>')
            best_sim, best_index = sim, self._index_of_row(row)
        hit = best_index >= 0 and not (best_sim < self.threshold)    # main.py:89
        return best_index, best_sim, hit

    def _top1(self, vec: np.ndarray) -> Tuple[int, float]:
        """(row relative to the head, similarity) of the best scanned row for one raw query.  ONE
        kernel: normalise the query + scan + top-1 (sqe_search_gemv, k = 1) -- replayed from a
        captured CUDA graph (H2D, kernel, D2H) when the cache state has been looked up before."""
        live, rows = self._scan()
        if self.use_graphs:
            g = self._graph
            key = (self._version, rows, live.data_ptr())
            stale = g is None or self._graph_key_of_graph != key
            if stale and self._graph_key != key:
                # first lookup of this cache state: launch eagerly.  The reference's handler
                # alternates get (miss) -> put (main.py:493, :547); capturing a graph per state
                # would cost more than it saves.  A second lookup of the same state captures.
                self._graph_key = key
                g = self._graph = None
            elif stale and time.monotonic() >= self._graph_retry_at:
                try:
                    g = self._graph = ops.SingleQueryGraph(live, rows, 1)
                    self._graph_key_of_graph = key
                except Exception as e:                       # capture not possible right now: eager, retry later
                    print(f"[GpuQueryCache] CUDA graph capture failed ({e}); eager launches for "
                          f"{self._GRAPH_RETRY_S:.0f} s")
                    self._graph_retry_at = time.monotonic() + self._GRAPH_RETRY_S
                    g = None
            elif stale:
                g = None
            if g is not None:
                s_, r_ = g.run(vec)
                return int(r_[0]), float(s_[0])
        buf, s, i = ops.packed_topk_out(self.device, 1, 1)
        ops.search_gemv(live, self._to_device(vec), 1, n=rows, out=(s, i))
        self._pinned_out[:12].copy_(buf, non_blocking=True)     # one 12-byte copy back
        torch.cuda.current_stream(self.device).synchronize()
        raw = self._pinned_out[:12].numpy()
        return int(raw[:8].view(np.int64)[0]), float(raw[8:12].view(np.float32)[0])

    _graph_key_of_graph = None

    def get(self, query_emb) -> Optional[str]:
        """lfu_cache_get, main.py:67-98."""
        with self._lock:
            if not self._live:                                       # main.py:70-71
                return None
            idx, _sim, hit = self.lookup(query_emb)
            if not hit:                                              # main.py:89-90
                return None
            e = self._entry_at_index(idx)
            e.freq = e.freq + 1                                      # main.py:94
            heapq.heappush(self._heap, (e.freq, -e.seq, e.row))
            if self.redis is not None:
                self.redis.lset(self.list_name, idx, self._entry_json(e))   # main.py:95
            return e.response

    def _entry_at_index(self, idx: int) -> _Entry:
        """The live entry with list index `idx`: row = head + idx + (tombstones in front of it)."""
        row = self._head + idx
        for d in self._dead_rows:                    # ascending; usually empty
            if d <= row:
                row += 1
            else:
                break
        return self._slot[row]

    def lookup_batch(self, queries: np.ndarray, path: int = 0
                     ) -> Tuple[np.ndarray, np.ndarray, np.ndarray]:
        """Streaming form (BASELINE config 5): host fp32 [B,1024] -> host
        (idx int32 [B], score fp32 [B], hit uint8 [B]).  Does not touch `freq`.  `idx` is the
        list index when the cache holds no tombstones (always the case after `bulk_load`), else the
        scanned row relative to the head (`index_of_scanned_row` converts)."""
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

    def index_of_scanned_row(self, rel_row: int) -> int:
        """List index of the entry at scanned row `rel_row` (as returned by `lookup_batch`); -1 for
        a tombstone."""
        row = self._head + int(rel_row)
        return -1 if rel_row < 0 or self._slot[row] is None else self._index_of_row(row)

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
        """fp32 CUDA queries [B,1024] (raw) -> CUDA (idx int32, score fp32, hit uint8) over the
        scanned rows.  `path`: 0 = choose (K2p when the cache keeps its int8 copy and B > 2, else
        as `sqe_cache_top1`), 1 = force the GEMV scan, 2 = force the 16-bit tensor-core scan,
        3 = force K2p."""
        live, rows = self._scan()
        q_dev = q_dev.contiguous()
        if (path == 3 or (path == 0 and self.prefilter and q_dev.shape[0] > 2 and
                          ops.k2p_pays(rows, q_dev.shape[0], self.dtype))) and rows > 0:
            if self._c8 is None:
                raise ValueError("this cache was built without prefilter=True")
            return ops.cache_top1_prefiltered(live, self._c8[self._head:], self._cm[self._head:], q_dev,
                                              self.threshold, n=rows, out=out)
        qn = ops.normalize_cast(q_dev, self.dtype)
        return ops.cache_top1(live, qn, self.threshold, path=path, n=rows, out=out)

    # ------------------------------------------------------------------- put
    def _write_row(self, row: int, src: torch.Tensor) -> None:
        """K1 (+ K1q) of fp32 device rows `src` into buffer rows [row, row + len)."""
        n = src.shape[0]
        ops.normalize_cast(src, self.dtype, out=self._buf[row: row + n])
        if self.prefilter:
            ops.quantize_rows(self._buf[row: row + n], out=(self._c8, self._cm), row0=row)

    def _zero_row(self, row: int) -> None:
        self._buf[row].zero_()
        if self.prefilter:
            self._c8[row].zero_()
            self._cm[row].zero_()

    def _pack(self) -> None:
        """The head reached row 0: pack the live rows to the top of the buffer (one gather)."""
        rows = list(self._live_rows())
        n = len(rows)
        new_head = self._cap - n
        if n:
            idx = torch.tensor(rows, dtype=torch.int64, device=self.device)
            self._buf[new_head:] = self._buf.index_select(0, idx)
            if self.prefilter:
                self._c8[new_head:] = self._c8.index_select(0, idx)
                self._cm[new_head:] = self._cm.index_select(0, idx)
        entries = [self._slot[r] for r in rows]
        self._slot = [None] * self._cap
        self._heap = []
        for j, e in enumerate(entries):
            e.row = new_head + j
            self._slot[e.row] = e
            self._heap.append((e.freq, -e.seq, e.row))
        heapq.heapify(self._heap)
        self._head = new_head
        self._dead_rows = []

    def _remove_least_frequent_item(self) -> None:
        """main.py:101-118: first entry with the minimal freq (strict '<'), i.e. among the entries
        with the smallest freq the one nearest to the head of the list (the newest)."""
        while self._heap:
            freq, nseq, row = heapq.heappop(self._heap)
            e = self._slot[row] if 0 <= row < self._cap else None
            if e is not None and e.freq == freq and e.seq == -nseq:
                break
        else:
            return
        if self.redis is not None:
            self.redis.lrem(self.list_name, 1, self._entry_json(e))  # main.py:117
        row = e.row
        self._slot[row] = None
        self._live -= 1
        ahead = row - self._head
        if ahead <= self._SLIDE_MAX and not self._dead_rows:
            # rows [head, row) slide to [head+1, row+1): list order is kept, nothing is left behind
            if ahead > 0:
                h = self._head
                self._buf[h + 1: row + 1] = self._buf[h: row].clone()
                if self.prefilter:
                    self._c8[h + 1: row + 1] = self._c8[h: row].clone()
                    self._cm[h + 1: row + 1] = self._cm[h: row].clone()
                for r in range(row, h, -1):
                    x = self._slot[r - 1]
                    self._slot[r] = x
                    if x is not None:
                        x.row = r
                        heapq.heappush(self._heap, (x.freq, -x.seq, r))
                self._slot[h] = None
            self._head += 1
        else:
            self._zero_row(row)                       # tombstone: similarity exactly 0.0, skipped on the host
            bisect.insort(self._dead_rows, row)

    def put(self, query_emb, response: str) -> None:
        """lfu_cache_put, main.py:121-128.  The reference stores the RAW embedding and
        normalises inside cosine_similarity at every lookup; storing the unit row once is the
        same cosine."""
        vec = self._row0(query_emb)
        if vec is None:
            return
        with self._lock, torch.cuda.device(self.device):
            if self._live >= self.max_items:                         # main.py:125-126
                self._remove_least_frequent_item()
            if self._live >= self.max_items:
                raise RuntimeError("cache buffer full (max_items reached with nothing evictable)")
            if self._head == 0:
                self._pack()
            self._head -= 1
            row = self._head
            self._write_row(row, self._to_device(vec))
            self._seq += 1
            e = _Entry(response, 1, vec.tolist() if self.keep_raw else None, self._seq, row)
            self._slot[row] = e
            self._live += 1
            self._version += 1
            heapq.heappush(self._heap, (1, -e.seq, row))
            if self.redis is not None:
                self.redis.lpush(self.list_name, self._entry_json(e))  # main.py:128

    def bulk_load(self, embeddings, responses=None) -> None:
        """Fill an empty cache with many entries at once (config-5 sized caches).
        Row i becomes list index i.  `embeddings`: host ndarray or CUDA fp32 tensor."""
        n = int(embeddings.shape[0])
        if self._live:
            raise RuntimeError("bulk_load needs an empty cache")
        if n > self.max_items:
            raise ValueError("more entries than max_items")
        with self._lock, torch.cuda.device(self.device):
            self._head = self._cap - n
            self._dead_rows = []
            step = 1 << 18
            for lo in range(0, n, step):
                hi = min(n, lo + step)
                blk = embeddings[lo:hi]
                if isinstance(blk, np.ndarray):
                    blk = torch.from_numpy(np.ascontiguousarray(blk, dtype=np.float32))
                blk = blk.to(self.device).contiguous()
                self._write_row(self._head + lo, blk)
            torch.cuda.current_stream(self.device).synchronize()
            self._slot = [None] * self._cap
            self._heap = []
            # list index i is NEWER than i + 1: sequence numbers fall with the index
            self._seq = n
            for i in range(n):
                e = _Entry(responses[i] if responses is not None else str(i), 1, None, n - i, self._head + i)
                self._slot[e.row] = e
                self._heap.append((1, -e.seq, e.row))
            heapq.heapify(self._heap)
            self._live = n
            self._version += 1

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
            self._heap = []
            for i, theirs in enumerate(parsed):
                mine = self._slot[self._head + i]
                mine.freq = theirs.get("freq", 1)                                # main.py:107
                if self.keep_raw:
                    mine.raw = theirs["embedding"]
                self._heap.append((mine.freq, -mine.seq, mine.row))
            heapq.heapify(self._heap)
        return len(parsed)

    # ------------------------------------------------------------- inspection
    def responses(self) -> List[str]:
        return [self._slot[r].response for r in self._live_rows()]

    def freqs(self) -> List[int]:
        return [self._slot[r].freq for r in self._live_rows()]

    def tombstones(self) -> int:
        return len(self._dead_rows)
