#!/usr/bin/env python
"""A/B of the one-query sharded step inside ONE torchrun job (box-to-box variance cancels):
exchange kernel vs exchange fused into the scan, ordinary vs overlapped launches, exact vs
prefiltered scan.  10M x 1024 bf16 over the ranks.
    python -m torch.distributed.run --nproc-per-node N ... scripts/multi_b1_ab.py [rows]"""
import json, os, sys
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import torch, torch.distributed as dist
import sqe_b200

rows_total = int(sys.argv[1]) if len(sys.argv) > 1 else 10_000_000
rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
local = int(os.environ.get("LOCAL_RANK", rank))
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
dist.init_process_group("nccl", device_id=dev)
BLK = 250_000
blo, bhi = sqe_b200.shard_bounds(rows_total // BLK, world, rank)
index = sqe_b200.GpuCorpusIndex(dtype="bf16", device=dev, keep_payload=False, prefilter=True)
index.reserve((bhi - blo) * BLK)
gen = torch.Generator(device=dev)
for blk in range(blo, bhi):
    gen.manual_seed(1234 + blk)
    index.add_device_rows(torch.randn((BLK, 1024), generator=gen, device=dev))
sh = sqe_b200.ShardedCorpusIndex(index)
sh.finalize()
q = torch.randn((1, 1024), generator=torch.Generator().manual_seed(99)).to(dev)
res = {}
for rep in range(2):
    for pre in (False, True):
        index.prefilter = pre
        for fused in (False, True):
            sh.fuse_small_batches = fused
            for ready in (False, True):
                if ready and not fused and world > 1:
                    continue
                for _ in range(10):
                    sh.search_device(q, 10, queries_ready=ready)
                dist.barrier(); torch.cuda.synchronize()
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                n = 300
                e0.record()
                for _ in range(n):
                    sh.search_device(q, 10, queries_ready=ready)
                e1.record()
                dist.barrier(); torch.cuda.synchronize()
                t = torch.tensor([e0.elapsed_time(e1) / n], dtype=torch.float64, device=dev)
                dist.all_reduce(t, op=dist.ReduceOp.MAX)
                key = f"{'K3p' if pre else 'K3'} {'fused' if fused else 'xchg-kernel'} {'overlapped' if ready else 'ordinary'}"
                res.setdefault(key, []).append(round(float(t.item()) * 1e3, 1))
if rank == 0:
    out = {"n_gpus": world, "rows_total": rows_total, "us_per_step": res,
           "qps": {k: round(1e6 / min(v), 1) for k, v in res.items()}}
    print(json.dumps(out))
dist.destroy_process_group()
