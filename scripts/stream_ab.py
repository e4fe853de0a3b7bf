#!/usr/bin/env python
"""A/B of the three ways to drive the sharded index (device-resident loop, one synchronous
search_batch per step, streaming search_batches), interleaved, under torchrun on N GPUs.
   torchrun --nproc-per-node N scripts/stream_ab.py [rows] [b] [k] [steps] [rounds]"""
import os, sys, time
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import numpy as np, torch, torch.distributed as dist
import sqe_b200
rank = int(os.environ.get("RANK", 0)); world = int(os.environ.get("WORLD_SIZE", 1)); lr = int(os.environ.get("LOCAL_RANK", 0))
torch.cuda.set_device(lr); dev = torch.device("cuda", lr)
if world > 1:
    dist.init_process_group("nccl", device_id=dev)
rows = int(sys.argv[1]) if len(sys.argv) > 1 else 10_000_000
b = int(sys.argv[2]) if len(sys.argv) > 2 else 1024
k = int(sys.argv[3]) if len(sys.argv) > 3 else 10
steps = int(sys.argv[4]) if len(sys.argv) > 4 else 20
rounds = int(sys.argv[5]) if len(sys.argv) > 5 else 3
depths = [int(x) for x in (sys.argv[6].split(",") if len(sys.argv) > 6 else ["2"])]
lo, hi = sqe_b200.shard_bounds(rows, world, rank)
local = sqe_b200.GpuCorpusIndex(dtype="bf16", device=dev, keep_payload=False)
local.reserve(hi - lo)
gen = torch.Generator(device=dev)
for s in range(lo, hi, 250_000):
    gen.manual_seed(s)
    local.add_device_rows(torch.randn((min(250_000, hi - s), 1024), generator=gen, device=dev))
sh = sqe_b200.ShardedCorpusIndex(local); sh.finalize()
qh = torch.randn((b, 1024)).pin_memory(); qd = qh.to(dev); qn = qh.numpy()

def timed(fn):
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    t0 = time.perf_counter(); fn(); torch.cuda.synchronize()
    t = torch.tensor([time.perf_counter() - t0], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item()) / steps * 1e3

def dev_loop():
    for _ in range(steps):
        sh.search_device(qd, k)
def sync_loop():
    for _ in range(steps):
        sh.search_batch(qn, k)
def stream_loop(depth):
    def f():
        for _ in sh.search_batches((qn for _ in range(steps)), k, depth=depth):
            pass
    return f
modes = [("device", dev_loop), ("sync", sync_loop)] + [(f"stream{d}", stream_loop(d)) for d in depths]
for name, fn in modes:
    timed(fn)
for r in range(rounds):
    res = {name: timed(fn) for name, fn in modes}
    if rank == 0:
        print(f"round {r}: " + "  ".join(f"{n_} {v:.3f} ms" for n_, v in res.items()), flush=True)
if world > 1:
    dist.destroy_process_group()
