"""Host logic of serving.MicroBatcher on the CPU: the GPU pipeline is replaced by a stand-in whose
"search result" is a function of the query, so routing (every requester gets exactly its own
rows, cut to its own k), the three client APIs (blocking, concurrent future, asyncio), error
behaviour and shutdown can be checked without a device.  The real pipeline is covered by
tests/test_gpu_parity.py::test_micro_batcher_coalesces_concurrent_requests."""
import asyncio
import queue
import threading

import numpy as np
import pytest

import sqe_b200
from sqe_b200 import ops, serving

DIM = 1024


class FakePipe:
    """StreamPipeline stand-in: row j of request i is 1000 * tag(i) + j, score = -j."""
    fail_next = False

    def __init__(self, dev, as_rows, out_bytes, launch, unpack, depth=2):
        self.depth = depth
        self._q = queue.Queue()
        self._free = queue.Queue()
        for _ in range(depth):
            self._free.put(1)

    def submit(self, batch, ctx=None):
        if FakePipe.fail_next:
            FakePipe.fail_next = False
            raise RuntimeError("injected launch failure")
        self._free.get()
        tags = batch[:, 0].astype(np.int64)
        k = int(ctx)
        rows = tags[:, None] * 1000 + np.arange(k)[None, :]
        scores = -np.arange(k, dtype=np.float32)[None, :].repeat(len(tags), 0)
        self._q.put((scores, rows))

    def collect(self):
        r = self._q.get()
        self._free.put(1)
        return r

    def in_flight(self):
        return self._q.qsize()


class FakeIndex:
    device = None
    strict = False
    keep_payload = False
    score_mode = "cosine"
    return_embedding = False
    _docs = []
    _as_rows = staticmethod(sqe_b200.GpuCorpusIndex._as_rows)
    hits_from_rows = sqe_b200.GpuCorpusIndex.hits_from_rows
    _source = sqe_b200.GpuCorpusIndex._source


@pytest.fixture
def batcher(monkeypatch):
    monkeypatch.setattr(ops, "StreamPipeline", FakePipe)
    made = []

    def make(**kw):
        idx = FakeIndex()
        idx.strict = kw.pop("strict", False)
        mb = serving.MicroBatcher(idx, **kw)
        made.append(mb)
        return mb
    yield make
    for mb in made:
        mb.close()


def _q(tag):
    q = np.zeros((1, DIM), dtype=np.float32)
    q[0, 0] = tag
    return q


def _expect(tag, k):
    return [({"doc_id": str(1000 * tag + j), "text": ""}, float(-j)) for j in range(k)]


def test_every_requester_gets_its_own_rows_through_all_three_apis(batcher):
    mb = batcher(max_batch=16, max_wait_s=2e-3, depth=2)
    out = {}

    def blocking(tag):
        out[("t", tag)] = mb.search(_q(tag), 1 + tag % 5)
    threads = [threading.Thread(target=blocking, args=(t,)) for t in range(40)]
    for t in threads:
        t.start()
    futs = {tag: mb.submit(_q(tag), 3) for tag in range(100, 130)}
    for t in threads:
        t.join()

    async def amain():
        res = await asyncio.gather(*[mb.asearch(_q(tag), 2 + tag % 3) for tag in range(200, 300)])
        wrapped = await asyncio.wrap_future(mb.submit(_q(77), 4))
        return res, wrapped
    ares, wrapped = asyncio.run(amain())
    for tag in range(40):
        assert out[("t", tag)] == _expect(tag, 1 + tag % 5)
    for tag, f in futs.items():
        assert f.result(timeout=10) == _expect(tag, 3)
    for tag, r in zip(range(200, 300), ares):
        assert r == _expect(tag, 2 + tag % 3)
    assert wrapped == _expect(77, 4)
    assert mb.requests == 40 + 30 + 100 + 1 and mb.batches < mb.requests       # it did coalesce


def test_blank_queries_errors_and_shutdown(batcher):
    mb = batcher(max_batch=8, max_wait_s=1e-3)
    assert mb.search(np.array([]), 3) == []                                      # main.py:350-351

    async def blank():
        return await mb.asearch(np.array([]), 3)
    assert asyncio.run(blank()) == []
    FakePipe.fail_next = True
    assert mb.search(_q(1), 3) == []                                             # main.py:371-373 -> []
    assert mb.search(_q(2), 2) == _expect(2, 2)                                  # and it keeps serving

    strict = batcher(max_batch=8, max_wait_s=1e-3, strict=True)
    FakePipe.fail_next = True
    with pytest.raises(RuntimeError, match="injected"):
        strict.search(_q(3), 3)

    async def strict_async():
        FakePipe.fail_next = True
        with pytest.raises(RuntimeError, match="injected"):
            await strict.asearch(_q(4), 3)
        return await strict.asearch(_q(5), 1)
    assert asyncio.run(strict_async()) == _expect(5, 1)
    mb.close()
    with pytest.raises(RuntimeError, match="closed"):
        mb.submit(_q(1), 1)


def test_a_cancelled_handler_does_not_break_its_batch(batcher):
    mb = batcher(max_batch=4, max_wait_s=20e-3)

    async def amain():
        a = mb.asubmit(_q(1), 2)
        b = mb.asubmit(_q(2), 2)
        a.cancel()
        return await b
    assert asyncio.run(amain()) == _expect(2, 2)
