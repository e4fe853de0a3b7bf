"""One rank of the 2+-GPU sharded parity test (launched by torchrun from
tests/test_gpu_parity_at_size.py::test_two_rank_sharded_search_vs_oracle, or by hand:

    python -m torch.distributed.run --nproc-per-node 2 --master-addr 127.0.0.1 \
        --master-port 29517 tests/multi_gpu_worker.py --exchange p2p

Every rank holds its row block of the corpus in a GpuCorpusIndex, answers through
ShardedCorpusIndex (local scan + exchange + merge) and checks ITS OWN merged result against the
oracle over the whole, unsharded corpus (tests/bigparity.py) -- not against a single-GPU run of
the CUDA path.  Exits non-zero on any mismatch."""
import argparse
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
sys.path.insert(0, os.path.dirname(HERE))

import sqe_b200                                            # noqa: E402
from bigparity import assert_topk_matches_at_size, oracle_candidates, stored_block_f32   # noqa: E402

DIM = 1024
BLOCK = 100_000


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--exchange", default="p2p", choices=["p2p", "nccl", "auto"])
    ap.add_argument("--rows", type=int, default=600_037)
    ap.add_argument("--dtype", default="bf16")
    ap.add_argument("--epochs", type=int, default=200)
    ap.add_argument("--stress", type=int, default=0,
                    help="protocol stress run: this many exchange epochs with skewed ranks (random host-side delays), "
                         "alternating the fused and the kernel exchange, batch shapes and scans; every result is "
                         "compared with the first result of the same request")
    args = ap.parse_args()
    rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
    local = int(os.environ.get("LOCAL_RANK", rank))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dist.init_process_group("nccl", device_id=dev)
    ops = sqe_b200.ops
    n, dtype = args.rows, args.dtype
    lo, hi = sqe_b200.shard_bounds(n, world, rank)

    g = torch.Generator().manual_seed(4242)
    q_raw = torch.randn((200, DIM), generator=g, dtype=torch.float32).to(dev)
    # ties ACROSS shards: the same raw row in the first and in the last shard (and twice in one)
    ties = {0: [n - 5, 3, n // world + 7], 150: [n - 1, 11]}
    index = sqe_b200.GpuCorpusIndex(dtype=dtype, device=dev, keep_payload=False)
    index.reserve(hi - lo)
    full = torch.empty((n, ops.ROW_ELEMS[dtype]), dtype=ops.TORCH_DTYPES[dtype], device=dev)   # oracle input only
    gen = torch.Generator(device=dev)
    for b0 in range(0, n, BLOCK):
        b1 = min(n, b0 + BLOCK)
        gen.manual_seed(900 + b0 // BLOCK)
        x = torch.randn((b1 - b0, DIM), generator=gen, device=dev)
        for qi, rows in ties.items():
            for r in rows:
                if b0 <= r < b1:
                    x[r - b0] = q_raw[qi] * 3.0
        ops.normalize_cast(x, dtype, out=full[b0:b1])
        a, z = max(b0, lo), min(b1, hi)
        if a < z:
            index.add_device_rows(x[a - b0: z - b0].contiguous())
    assert torch.equal(index.shard.view(torch.int16), full[lo:hi].view(torch.int16)), "shard != K1 of its rows"
    sharded = sqe_b200.ShardedCorpusIndex(index, exchange=args.exchange)
    sharded.finalize()
    assert sharded.row_offset == lo and sharded.total_rows == n

    Q = ops.normalize_cast(q_raw, dtype)
    q_st = stored_block_f32(Q, dtype, 0, Q.shape[0])
    report = []

    def check(b, k, tol, label):
        s, i = sharded.search_device(q_raw[:b].contiguous(), k)
        torch.cuda.synchronize()
        exc, worst = assert_topk_matches_at_size(s.cpu().numpy(), i.cpu().numpy(), full, dtype, n, q_st[:b], k,
                                                 score_tol=tol, tie_eps=1e-6)
        report.append(f"{label}: b={b} k={k} worst={worst:.1e} excused={exc}")
        return s, i

    s1, i1 = check(1, 10, 2e-6, "K3 with the exchange fused into the scan")
    assert i1[0, :3].tolist() == sorted(ties[0]), i1[0].tolist()
    sharded.fuse_small_batches = False
    s1b, i1b = check(1, 10, 2e-6, "K3 + exchange kernel")
    sharded.fuse_small_batches = True
    assert torch.equal(i1, i1b) and torch.equal(s1.view(torch.int32), s1b.view(torch.int32))
    sb, ib = check(200, 100, 1e-5, "K2 (R=4) + exchange")
    assert ib[0, :3].tolist() == sorted(ties[0]) and ib[150, :2].tolist() == sorted(ties[150])
    check(130, 10, 1e-5, "K2 + exchange")
    check(8, 33, 2e-6 if dtype == "fp32" else 1e-5, "small batch + exchange")
    index.enable_prefilter()
    s2, i2 = check(2, 10, 2e-6, "K3p + exchange")
    assert torch.equal(i2[:1], i1) and torch.equal(s2[:1].view(torch.int32), s1.view(torch.int32))
    # many epochs back to back (buffer parity, flags, no host sync in between): identical every time
    # -- with ordinary launches, then overlapped ones (the queries are resident: SQE_FLAG_QUERIES_READY),
    # with and without the int8 prefilter
    q4 = [q_raw[e: e + 1].contiguous() for e in range(4)]
    torch.cuda.synchronize()
    for pre in (True, False):
        index.prefilter = pre
        ref = None
        for ready in (False, True):
            outs = [sharded.search_device(q4[e % 4], 10, queries_ready=ready) for e in range(args.epochs)]
            torch.cuda.synchronize()
            for e in range(4, args.epochs):
                assert torch.equal(outs[e][1], outs[e % 4][1]) and torch.equal(outs[e][0], outs[e % 4][0]), (pre, ready, e)
            if ref is not None:
                for e in range(4):
                    assert torch.equal(outs[e][1], ref[e][1]) and torch.equal(outs[e][0], ref[e][0]), (pre, e)
            ref = outs
        assert torch.equal(ref[0][1], i1) and torch.equal(ref[0][0].view(torch.int32), s1.view(torch.int32))
    index.prefilter = True
    if args.stress:
        # skewed ranks: every rank sleeps for its own random times between launches, so the ranks enter
        # the exchanges out of step (a rank may be a full call ahead of a peer: buffer parity, epoch flags)
        import time as _time
        rng = np.random.default_rng(1000 + rank)
        shapes = [(1, 10, True), (1, 10, False), (4, 10, True), (2, 10, True), (1, 100, True), (130, 10, True), (2, 33, False)]
        first = {}
        t0 = _time.time()
        for e in range(args.stress):
            b_, k_, fused = shapes[e % len(shapes)]
            sharded.fuse_small_batches = fused
            index.prefilter = (e // len(shapes)) % 2 == 0
            if rng.random() < 0.3:
                _time.sleep(float(rng.uniform(0, 3e-4)) * (1 + 3 * (rank == (e // 97) % world)))
            out = sharded.search_device(q_raw[:b_].contiguous(), k_, queries_ready=bool(e % 2))
            # one-query scans are bit-identical with and without the prefilter; two queries and batches take the
            # 16-bit tensor kernel without it (agrees to 1e-5, not bit for bit): compare like with like
            key = (b_, k_, index.prefilter if b_ > 1 else None)
            if key not in first:
                torch.cuda.synchronize()
                first[key] = (out[0].clone(), out[1].clone())
            elif e % 7 == 0 or e > args.stress - 50:
                torch.cuda.synchronize()
                assert torch.equal(out[1], first[key][1]) and torch.equal(out[0], first[key][0]), (e, key)
        torch.cuda.synchronize()
        sharded.fuse_small_batches = True
        index.prefilter = True
        report.append(f"stress: {args.stress} exchange epochs with skewed ranks in {_time.time() - t0:.1f} s, "
                      f"{len(shapes)} request shapes, fused / kernel exchange, exact / prefiltered: all identical")
    report.append(f"{4 * args.epochs} back-to-back one-query steps (ordinary + overlapped launches, exact + prefiltered): identical")
    assert sharded.exchange == ("nccl" if args.exchange == "nccl" else sharded.exchange)
    dist.barrier()
    print(f"[rank {rank}] exchange={sharded.exchange} " + "; ".join(report) + " -- rank ok", flush=True)
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
