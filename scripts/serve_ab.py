#!/usr/bin/env python
"""A/B of MicroBatcher settings under asyncio clients (one event-loop thread, the reference's
handler model): depth (1 = one batcher thread, 2 = launcher + delivery threads), max_batch and
the interpreter's thread switch interval.   python scripts/serve_ab.py [rows] [clients]"""
import asyncio, os, sys, time
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import numpy as np, torch
import sqe_b200
rows = int(sys.argv[1]) if len(sys.argv) > 1 else 10_000_000
clients = int(sys.argv[2]) if len(sys.argv) > 2 else 256
per_client, k = 8, 10
dev = torch.device("cuda", 0)
index = sqe_b200.GpuCorpusIndex(dtype="bf16", device=dev, keep_payload=False)
index.reserve(rows)
gen = torch.Generator(device=dev)
for s in range(0, rows, 250_000):
    gen.manual_seed(s)
    index.add_device_rows(torch.randn((min(250_000, rows - s), 1024), generator=gen, device=dev))
qs = np.random.default_rng(3).standard_normal((clients, per_client, 1, 1024)).astype(np.float32)

def run(mb):
    lat = []
    async def aclient(c):
        for r in range(per_client):
            t0 = time.perf_counter()
            await asyncio.wrap_future(mb.submit(qs[c, r], k))
            lat.append(time.perf_counter() - t0)
    async def amain():
        await asyncio.gather(*[aclient(c) for c in range(clients)])
    t0 = time.perf_counter(); asyncio.run(amain()); dt = time.perf_counter() - t0
    lat.sort()
    return clients * per_client / dt, lat[len(lat) // 2] * 1e3, lat[int(len(lat) * 0.99)] * 1e3

for interval in (5e-3, 2e-4):
    sys.setswitchinterval(interval)
    for depth, mbs in ((1, 256), (1, 128), (2, 128), (2, 64), (3, 64)):
        mb = sqe_b200.MicroBatcher(index, max_batch=mbs, max_wait_s=500e-6, depth=depth)
        run(mb)
        b0, r0 = mb.batches, mb.requests
        best = max((run(mb) for _ in range(3)), key=lambda t: t[0])
        print(f"switchinterval {interval*1e3:.1f} ms depth {depth} max_batch {mbs}: {best[0]:.0f} req/s, "
              f"p50 {best[1]:.1f} ms, p99 {best[2]:.1f} ms, mean batch {(mb.requests-r0)/max(mb.batches-b0,1):.0f}", flush=True)
        mb.close()
