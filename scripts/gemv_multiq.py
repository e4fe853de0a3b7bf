#!/usr/bin/env python
"""K3 with several queries: time vs batch on an fp32 shard (the queries share rows through L2)."""
import os, sys
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import torch, sqe_b200
from sqe_b200 import ops
dev = torch.device("cuda", 0)
rows = 1_000_000
D = ops.normalize_cast(torch.randn((rows, 1024), device=dev), "fp32")
for nq in (1, 2, 4, 8, 16, 64):
    Q = ops.normalize_cast(torch.randn((nq, 1024), device=dev), "fp32")
    for _ in range(3): ops.topk_gemv(D, Q, 10)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(10): ops.topk_gemv(D, Q, 10)
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 10
    print(f"fp32 1M rows, nq={nq:3d}: {ms:8.3f} ms  = {ms / 0.585:5.2f} single-query passes, {nq / ms * 1e3:9.0f} q/s")
