#!/usr/bin/env python
"""Small end-to-end pass over every kernel for compute-sanitizer (memcheck / racecheck):
   compute-sanitizer --tool memcheck python scripts/sanitize_small.py"""
import os, sys
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import numpy as np, torch
import sqe_b200
from sqe_b200 import ops
nat = sqe_b200._native
dev = torch.device("cuda", 0)
g = torch.Generator(device=dev); g.manual_seed(0)
x = torch.randn((3001, 1024), generator=g, device=dev)
for dt in ("bf16", "fp16", "fp32"):
    D = ops.normalize_cast(x, dt)
    q = torch.randn((130, 1024), generator=g, device=dev)
    Q = ops.normalize_cast(q, dt)
    for k in (1, 10, 100):
        s1, i1 = ops.topk_gemv(D, Q[:3], k)
        s2, i2 = ops.search_gemv(D, q[:2], k)
        if dt != "fp32":
            for cg in (1, 2):
                nat.tuning_set(nat.SQE_TUNE_K2_CTA_GROUP, cg)
                sb, ib = ops.topk_batched(D, Q, min(k, 128))
                sb, ib = ops.topk_batched(D, Q[:5], min(k, 128), n=257)
            nat.tuning_set(nat.SQE_TUNE_K2_CTA_GROUP, 0)
    idx, sc, hit = ops.cache_top1(D, Q[:9], 0.96, path=0)
    idx, sc, hit = ops.cache_top1(D, Q[:9], 0.96, path=1)
s = torch.sort(torch.randn((4, 33, 10), generator=g, device=dev), dim=2, descending=True).values.contiguous()
i = torch.randperm(4 * 33 * 10, generator=g, device=dev).view(4, 33, 10)
ops.merge_topk(s, i, 10)
buf = torch.zeros(ops.exchange_buffer_bytes(1, 33 * 10), dtype=torch.uint8, device=dev)
ops.exchange_merge(s[0], i[0], 10, 0, [buf.data_ptr()], 33 * 10, 1)
torch.cuda.synchronize()
print("sanitize_small done")
