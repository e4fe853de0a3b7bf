"""Callers and data formats on either side of the path (SURVEY.md §8f "next" rows).

* `UserIndexRegistry` -- the per-user index namespaces of the upload micro-service
  (/root/reference/app/embedding_gen.py:83-122 `init_user_index`, :196-257
  `bulk_index_embeddings`): one `GpuCorpusIndex` per `"{BASE}-{user_id}"`.
* `MicroBatcher` -- handler-level micro-batching: the reference's handlers issue ONE query
  per request (app/main.py:499, :684); under concurrent load those single queries are
  coalesced into one batched search, i.e. one tensor-core launch (K2) instead of N streaming
  passes over the shard (K3), and every request still gets exactly its own result.
* `group_hits_by_doc` / `build_context_text` -- the step right after the path
  (app/main.py:500-513): chunks of the same document are concatenated in hit order.
"""
from __future__ import annotations

import os
import threading
import time
from concurrent.futures import Future
from typing import Callable, Dict, List, Optional, Sequence, Tuple

import numpy as np
import torch

from . import ops
from .corpus import GpuCorpusIndex

BASE_OPENSEARCH_INDEX_NAME = os.getenv("OPENSEARCH_INDEX_NAME", "")   # embedding_gen.py:38


# --------------------------------------------------------------------- per-user indices
class UserIndexRegistry:
    """Drop-in for the module-level functions of embedding_gen.py."""

    def __init__(self, base_name: str = BASE_OPENSEARCH_INDEX_NAME, **index_kwargs):
        self.base_name = base_name
        self.index_kwargs = index_kwargs
        self._lock = threading.Lock()
        self._indices: Dict[str, GpuCorpusIndex] = {}

    def index_name(self, user_id: str) -> str:
        return f"{self.base_name}-{user_id}"                      # embedding_gen.py:91, :211

    def init_user_index(self, user_id: str) -> GpuCorpusIndex:
        """embedding_gen.py:83-122: create the user's index if it does not exist yet."""
        name = self.index_name(user_id)
        with self._lock:
            idx = self._indices.get(name)
            if idx is None:
                idx = GpuCorpusIndex(None, name, **self.index_kwargs)
                self._indices[name] = idx
            else:
                print(f"[INFO] Index '{name}' already exists.")    # embedding_gen.py:93
            return idx

    def bulk_index_embeddings(self, user_id: str, doc_id: str, embeddings: np.ndarray,
                              chunks: Sequence[str]) -> None:
        """embedding_gen.py:196-257: all chunks of one document into the user's index,
        `_id = f"{doc_id}_{chunk_index}"`; rows are normalised on the GPU (K1)."""
        if embeddings is None or getattr(embeddings, "size", 0) == 0:
            print("[ERROR] Missing OpenSearch client or embeddings => cannot index.")   # :207-209
            return
        idx = self.init_user_index(user_id)
        idx.add_document_chunks(doc_id, embeddings, chunks)

    def get(self, user_id: str) -> Optional[GpuCorpusIndex]:
        return self._indices.get(self.index_name(user_id))

    def search(self, user_id: str, query_emb: np.ndarray, k: int = 3):
        idx = self.get(user_id)
        return [] if idx is None else idx.search(query_emb, k)


# ------------------------------------------------------------------- micro-batching
class _LoopFuture:
    """A request that waits on an event loop: its asyncio future and the loop that owns it."""
    __slots__ = ("loop", "afut")

    def __init__(self, loop, afut):
        self.loop, self.afut = loop, afut

    def done(self) -> bool:
        return self.afut.done()


def _resolve_on_loop(group) -> None:
    for afut, value, is_exc in group:                                # runs ON the loop thread
        if afut.done():                                              # cancelled by its handler
            continue
        if is_exc:
            afut.set_exception(value)
        else:
            afut.set_result(value)


class MicroBatcher:
    """Coalesce concurrent single-query `search` calls into batched launches.

    `search(query_emb, k)` blocks like `GpuCorpusIndex.search` and returns the same list;
    `submit(query_emb, k)` returns a `concurrent.futures.Future` (wrap it with
    `asyncio.wrap_future` inside the FastAPI handlers).  A launcher thread drains the queue:
    it waits at most `max_wait_s` after the first request for more to arrive (or until
    `max_batch` are queued) and enqueues ONE batched search with the largest k requested (copies
    and kernels are asynchronous, ops.StreamPipeline).  A delivery thread waits for the oldest
    batch in flight and slices each request's rows out of it (a top-k list is a prefix of a longer
    top-k list) -- so the GPU scores batch j+1 while the results of batch j are being handed out
    (`depth` batches in flight).

    `await mb.asearch(query_emb, k)` is the form for `async def` handlers on an event loop
    (main.py:587, :650): the request carries an asyncio future of the CALLER's loop and the delivery
    thread resolves all requests of a batch with ONE `call_soon_threadsafe` per loop -- one
    loop wake-up per batch instead of one per request (`asyncio.wrap_future(mb.submit(...))` costs a
    concurrent future, a chained asyncio future and a self-pipe write per request)."""

    def __init__(self, index: GpuCorpusIndex, max_batch: int = 256, max_wait_s: float = 200e-6,
                 depth: int = 2, encoder=None):
        """`encoder` (a `GpuEmbeddingEncoder` on the index's device): `submit_text` / `asearch_text`
        then take the QUERY TEXT (main.py:676 `embed_query` + :684 `os_search` in one request): the
        batch's token lists go through one packed encoder pass on the compute stream and its CLS
        embeddings are searched where they are -- no embedding ever visits the host."""
        import queue
        self.index = index
        self.encoder = encoder
        self.max_batch = int(max_batch)
        self.max_wait_s = float(max_wait_s)
        self._cv = threading.Condition()
        self._queue: List[Tuple[np.ndarray, int, Future]] = []
        self._stop = False
        self.batches = 0               # batched launches issued
        self.requests = 0              # requests served

        def launch(qd, buf, k):
            b = qd.shape[0]
            index.search_device(qd, k, out=(buf[b * k * 8:].view(torch.float32).view(b, k),
                                            buf[: b * k * 8].view(torch.int64).view(b, k)))

        def unpack(arr, b, k):
            return (arr[b * k * 8:].view(np.float32).reshape(b, k).copy(),
                    arr[: b * k * 8].view(np.int64).reshape(b, k).copy())

        self._pipe = ops.StreamPipeline(index.device, lambda q: q, lambda b, k: b * k * 12, launch, unpack, depth)
        self._handoff: "queue.Queue" = queue.Queue()
        self._thread = threading.Thread(target=self._run, name="sqe-microbatcher", daemon=True)
        self._deliverer = threading.Thread(target=self._deliver, name="sqe-microbatcher-out", daemon=True)
        self._thread.start()
        if depth > 1:
            self._deliverer.start()

    # -- client side
    def submit(self, query_emb: np.ndarray, k: int = 3) -> Future:
        fut: Future = Future()
        if query_emb is None or getattr(query_emb, "size", 0) == 0:     # main.py:350-351
            fut.set_result([])
            return fut
        q = self.index._as_rows(query_emb)[:1]
        with self._cv:
            if self._stop:
                raise RuntimeError("MicroBatcher is closed")
            self._queue.append((q, int(k), fut))
            self._cv.notify()
        return fut

    def search(self, query_emb: np.ndarray, k: int = 3):
        return self.submit(query_emb, k).result()

    # text requests: the queue entry carries the token ids instead of an embedding
    def _ids_of(self, query):
        if self.encoder is None:
            raise RuntimeError("this MicroBatcher was built without an encoder")
        if isinstance(query, str):
            if not query.strip():                                        # main.py:176-177 -> [] downstream
                return None
            return self.encoder.tok.encode(query)
        return list(query) if len(query) else None

    def submit_text(self, query, k: int = 3) -> Future:
        """`query`: the text (tokenised here, on the caller's thread) or its token ids."""
        fut: Future = Future()
        ids = self._ids_of(query)
        if ids is None:
            fut.set_result([])
            return fut
        with self._cv:
            if self._stop:
                raise RuntimeError("MicroBatcher is closed")
            self._queue.append((ids, int(k), fut))
            self._cv.notify()
        return fut

    def search_text(self, query, k: int = 3):
        return self.submit_text(query, k).result()

    async def asearch_text(self, query, k: int = 3):
        import asyncio
        loop = asyncio.get_running_loop()
        afut = loop.create_future()
        ids = self._ids_of(query)
        if ids is None:
            afut.set_result([])
            return await afut
        with self._cv:
            if self._stop:
                raise RuntimeError("MicroBatcher is closed")
            self._queue.append((ids, int(k), _LoopFuture(loop, afut)))
            self._cv.notify()
        return await afut

    def asubmit(self, query_emb: np.ndarray, k: int = 3):
        """Called ON an event loop: returns an asyncio future of that loop (await it)."""
        import asyncio
        loop = asyncio.get_running_loop()
        afut = loop.create_future()
        if query_emb is None or getattr(query_emb, "size", 0) == 0:     # main.py:350-351
            afut.set_result([])
            return afut
        q = self.index._as_rows(query_emb)[:1]
        with self._cv:
            if self._stop:
                raise RuntimeError("MicroBatcher is closed")
            self._queue.append((q, int(k), _LoopFuture(loop, afut)))
            self._cv.notify()
        return afut

    async def asearch(self, query_emb: np.ndarray, k: int = 3):
        return await self.asubmit(query_emb, k)

    def close(self) -> None:
        with self._cv:
            self._stop = True
            self._cv.notify()
        self._thread.join()
        if self._pipe.depth > 1:
            self._deliverer.join()

    # -- server side
    def _take(self) -> List[Tuple[np.ndarray, int, Future]]:
        with self._cv:
            while not self._queue and not self._stop:
                self._cv.wait()
            if not self._queue:
                return []
            deadline = time.perf_counter() + self.max_wait_s
            while len(self._queue) < self.max_batch and not self._stop:
                left = deadline - time.perf_counter()
                if left <= 0:
                    break
                self._cv.wait(left)
            batch, self._queue = self._queue[: self.max_batch], self._queue[self.max_batch:]
            return batch

    def _fail(self, batch, e: Exception) -> None:
        strict = self.index.strict
        self._resolve([(fut, e if strict else [], strict) for _, _, fut in batch])   # main.py:371-373 -> []

    @staticmethod
    def _resolve(items) -> None:
        """items: (future, value, is_exception).  Thread futures are completed here; asyncio
        futures are handed to their loop in ONE call per loop."""
        per_loop = {}
        for fut, value, is_exc in items:
            if isinstance(fut, _LoopFuture):
                per_loop.setdefault(fut.loop, []).append((fut.afut, value, is_exc))
            elif is_exc:
                fut.set_exception(value)
            else:
                fut.set_result(value)
        for loop, group in per_loop.items():
            try:
                loop.call_soon_threadsafe(_resolve_on_loop, group)
            except RuntimeError:                                     # that loop is closed: nobody waits
                pass

    def _run(self) -> None:
        while True:
            batch = self._take()
            if not batch:
                if self._stop:
                    self._handoff.put(None)
                    return
                continue
            # embeddings and texts are separate launches (a batch is usually all of one kind)
            groups = [[b for b in batch if isinstance(b[0], np.ndarray)], [b for b in batch if not isinstance(b[0], np.ndarray)]]
            for is_text, group in enumerate(groups):
                if not group:
                    continue
                try:
                    kmax = max(b[1] for b in group)
                    if is_text:
                        seqs = [b[0] for b in group]
                        self._pipe.submit_device(lambda: self.encoder.forward_ids(seqs), len(seqs), kmax)
                    else:
                        self._pipe.submit(np.concatenate([b[0] for b in group], axis=0), kmax)   # blocks while `depth` are in flight
                    self.batches += 1
                    self.requests += len(group)
                    if self._pipe.depth > 1:
                        self._handoff.put(group)
                    else:
                        self._deliver_one(group)                     # depth 1: one thread does it all
                except Exception as e:
                    self._fail([b for b in group if not b[2].done()], e)

    def _deliver_one(self, batch) -> None:
        scores, rows = self._pipe.collect()
        scores, rows = scores.tolist(), rows.tolist()                # one conversion per batch
        hits = self.index.hits_from_rows
        self._resolve([(fut, hits(scores[i][:k], rows[i][:k]), False) for i, (_, k, fut) in enumerate(batch)])

    def _deliver(self) -> None:
        while True:
            batch = self._handoff.get()
            if batch is None:
                return
            try:
                self._deliver_one(batch)
            except Exception as e:
                self._fail([b for b in batch if not b[2].done()], e)


# ------------------------------------------------------- the step right after the path
def group_hits_by_doc(hits) -> Dict[str, str]:
    """What main.py:500-507 builds: for every doc_id (first-seen order) the texts of its chunks
    in hit order, joined by newlines."""
    per_doc: Dict[str, List[str]] = {}
    for source, _score in hits:
        per_doc.setdefault(source["doc_id"], []).append(source["text"])
    return {doc_id: "\n".join(texts) for doc_id, texts in per_doc.items()}


def build_context_text(hits) -> str:
    """The prompt context of main.py:509-513: one '--- Document ID: x ---' block per document."""
    return "".join(f"--- Document ID: {doc_id} ---\n{text}\n\n"
                   for doc_id, text in group_hits_by_doc(hits).items())
