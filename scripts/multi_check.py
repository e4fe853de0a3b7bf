#!/usr/bin/env python
"""N-rank check of the sharded search (run under torchrun on N GPUs): the fused P2P exchange
and the NCCL exchange must give identical results, equal to the single-shard result; then
time both.   torchrun --nproc-per-node N scripts/multi_check.py"""
import os, sys
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import numpy as np, torch, torch.distributed as dist
import sqe_b200
rank = int(os.environ["RANK"]); world = int(os.environ["WORLD_SIZE"]); lr = int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(lr); dev = torch.device("cuda", lr)
dist.init_process_group("nccl", device_id=dev)
n = 400_000
gen = torch.Generator(device=dev); gen.manual_seed(5)
full = torch.randn((n, 1024), generator=gen, device=dev)          # same on every rank (same seed)
full[n - 7] = full[11]                                            # tie across shards
lo, hi = sqe_b200.shard_bounds(n, world, rank)
# SQE_PREFILTER=1: the shards answer one- and two-query searches through the int8-prefiltered scan
# (K3p); the single-shard reference below stays on the exact scan
local = sqe_b200.GpuCorpusIndex(dtype="bf16", device=dev, keep_payload=False,
                                prefilter=bool(os.environ.get("SQE_PREFILTER")))
local.add_device_rows(full[lo:hi])
whole = sqe_b200.GpuCorpusIndex(dtype="bf16", device=dev, keep_payload=False)
whole.add_device_rows(full)
for b, k in [(1, 10), (2, 40), (64, 10), (1024, 10), (256, 100)]:
    q = torch.randn((b, 1024), generator=gen, device=dev); q[0] = full[11] * 2
    want_s, want_i = whole.search_device(q, k)
    res = {}
    for mode in ("p2p", "nccl"):
        sh = sqe_b200.ShardedCorpusIndex(local, exchange=mode); sh.finalize()
        for rep in range(3):
            s, i = sh.search_device(q, k)
        torch.cuda.synchronize()
        assert sh.exchange == mode, (sh.exchange, mode)
        assert torch.equal(i, want_i), (mode, b, k, (i != want_i).sum().item())
        assert torch.allclose(s, want_s, atol=1e-5)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        dist.barrier(); torch.cuda.synchronize()
        e0.record()
        for _ in range(50):
            sh.search_device(q, k)
        e1.record(); torch.cuda.synchronize()
        res[mode] = e0.elapsed_time(e1) / 50
        # the exchange alone, on fixed local lists
        ls, li = local.search_device(q, k, idx_offset=sh.row_offset)
        for _ in range(5):
            sh.exchange_lists(ls, li, k)
        dist.barrier(); torch.cuda.synchronize()
        e0.record()
        for _ in range(200):
            sh.exchange_lists(ls, li, k)
        e1.record(); torch.cuda.synchronize()
        res[mode + "_x"] = e0.elapsed_time(e1) / 200
    assert want_i[0, 0].item() == 11 and want_i[0, 1].item() == n - 7
    # the streaming host API over the sharded index: every batch identical to the single shard
    sh = sqe_b200.ShardedCorpusIndex(local, exchange="p2p"); sh.finalize()
    qh = q.cpu().numpy()
    outs = list(sh.search_batches([qh, qh[: max(1, b // 2)], qh, qh], k))
    wi = want_i.cpu().numpy()
    assert len(outs) == 4 and all(np.array_equal(o[1], wi[: o[1].shape[0]]) for o in outs), "streaming mismatch"
    if rank == 0:
        print(f"world={world} b={b} k={k}: identical to single shard; step p2p {res['p2p']*1e3:.1f} us, nccl {res['nccl']*1e3:.1f} us; "
              f"exchange alone p2p {res['p2p_x']*1e3:.1f} us, nccl {res['nccl_x']*1e3:.1f} us", flush=True)
dist.destroy_process_group()
