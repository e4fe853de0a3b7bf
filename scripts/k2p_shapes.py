#!/usr/bin/env python
"""K2p vs K2 over shard sizes / batch sizes on ONE corpus (prefixes of 10M rows): ms per call, pairs logged,
first query checked against the exact scan."""
import os, sys
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import numpy as np
import torch
import sqe_b200
from sqe_b200 import ops
nat = sqe_b200._native
dev = torch.device("cuda", 0)
ROWS = int(sys.argv[1]) if len(sys.argv) > 1 else 10_000_000
D = torch.empty((ROWS, 1024), dtype=torch.bfloat16, device=dev)
gen = torch.Generator(device=dev)
for lo in range(0, ROWS, 250_000):
    gen.manual_seed(1234 + lo // 250_000)
    ops.normalize_cast(torch.randn((min(250_000, ROWS - lo), 1024), generator=gen, device=dev), "bf16", out=D[lo:lo + 250_000])
d8, meta = ops.quantize_rows(D)
Q = torch.randn((1024, 1024), generator=torch.Generator().manual_seed(99)).to(dev)
resc = torch.zeros((1024,), dtype=torch.int32, device=dev)


def timeit(fn, n=10):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n


for rep in range(2):
    for rows, b, k in [(10_000_000, 1024, 10), (5_000_000, 1024, 10), (2_500_000, 1024, 10), (1_250_000, 1024, 10),
                       (10_000_000, 128, 10), (1_000_000, 64, 1), (2_500_000, 256, 10)]:
        if rows > ROWS:
            continue
        q = Q[:b]
        qn = ops.normalize_cast(q, "bf16")
        tp = timeit(lambda: ops.search_batched_prefiltered(D[:rows], d8[:rows], meta[:rows], q, k, rescored=resc))
        exact = int(np.median(resc[:b].cpu().numpy()))
        t2 = timeit(lambda: ops.topk_batched(D[:rows], qn, k))
        tp2 = timeit(lambda: ops.search_batched_prefiltered(D[:rows], d8[:rows], meta[:rows], q, k, rescored=resc))
        want_s, want_i = ops.search_gemv(D[:rows], q[:1], k)
        s, i = ops.search_batched_prefiltered(D[:rows], d8[:rows], meta[:rows], q, k)
        ok = bool(torch.equal(s[:1], want_s) and torch.equal(i[:1], want_i))
        print(f"rows={rows} b={b} k={k}: K2p {tp:.3f} / {tp2:.3f} ms (before / after the K2 block), K2 {t2:.3f} ms, "
              f"rows scored exactly {exact}{'' if ok else ' MISMATCH'}", flush=True)
