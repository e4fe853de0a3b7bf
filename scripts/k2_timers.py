#!/usr/bin/env python
"""Role timers of the tensor-core kernel (diagnostics): where do the MMA issuer, the TMA
producer and the epilogue spend their cycles?   python scripts/k2_timers.py [rows] [b] [cg]"""
import os, sys
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import torch
import sqe_b200
from sqe_b200 import ops
nat = sqe_b200._native

rows = int(sys.argv[1]) if len(sys.argv) > 1 else 2_000_000
b = int(sys.argv[2]) if len(sys.argv) > 2 else 1024
K = int(sys.argv[4]) if len(sys.argv) > 4 else 10
DT = sys.argv[5] if len(sys.argv) > 5 else "bf16"
if len(sys.argv) > 6:
    nat.tuning_set(nat.SQE_TUNE_K2_WINDOW, int(sys.argv[6]))
TDT = {"bf16": torch.bfloat16, "fp16": torch.float16}[DT]
dev = torch.device("cuda", 0)
D = torch.empty((rows, 1024), dtype=TDT, device=dev)
gen = torch.Generator(device=dev)
for lo in range(0, rows, 250_000):
    gen.manual_seed(lo)
    x = torch.randn((min(250_000, rows - lo), 1024), generator=gen, device=dev)
    ops.normalize_cast(x, DT, out=D[lo:lo + x.shape[0]])
Q = ops.normalize_cast(torch.randn((b, 1024), generator=gen, device=dev), DT)
names = ["prod_total", "prod_wait_empty", "mma_total", "mma_wait_full", "mma_wait_tempty", "epi_total", "epi_wait_tfull", "epi_flush"]
import itertools
for cg, mode in itertools.product(([int(sys.argv[3])] if len(sys.argv) > 3 and int(sys.argv[3]) else [1, 2]), (0,)):
    nat.tuning_set(nat.SQE_TUNE_K2_CTA_GROUP, cg)
    nat.tuning_set(nat.SQE_TUNE_K2_EPILOGUE_MODE, mode)
    for _ in range(3):
        ops.topk_batched(D, Q, K)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(5):
        ops.topk_batched(D, Q, K)
    e1.record()
    torch.cuda.synchronize()
    print(f"epilogue_mode={mode} cta_group={cg}: {e0.elapsed_time(e1) / 5:.3f} ms per call without timers "
          f"-> {2 * rows * b * 1024 / (e0.elapsed_time(e1) / 5 * 1e-3) / 1e12:.0f} TFLOP/s")
    buf = torch.zeros((160 * 40 + 64 * 4,), dtype=torch.int64, device=dev)
    nat.load().sqe_debug_k2_timers(buf.data_ptr())
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    ops.topk_batched(D, Q, K)
    e1.record()
    torch.cuda.synchronize()
    nat.load().sqe_debug_k2_timers(None)
    grid = 148
    t = buf[: grid * 40].view(grid, 40).cpu().double()
    print(f"epilogue_mode={mode} cta_group={cg} rows={rows} b={b}: {e0.elapsed_time(e1):.3f} ms (timed with the timers on)")
    def stat(nm, col):
        col = col[col > 0]
        if len(col):
            print(f"  {nm:22s} n={len(col):3d} mean={col.mean().item():11.0f} min={col.min().item():11.0f} max={col.max().item():11.0f}")
    for j, nm in enumerate(["prod_total", "prod_wait_empty", "mma_total", "mma_wait_full", "mma_wait_tempty"]):
        stat(nm, t[:, j])
    stat("prod_window_wait", t[:, 32])
    stat("prod_window_polls", t[:, 33])
    for w in range(4):
        for j, nm in enumerate(["total", "wait_tfull", "flush", "n_slow_strips", "n_list_merges", "tmem_ld"]):
            stat(f"epi_warp{w + 2}_{nm}", t[:, 8 + 6 * w + j])
    tr = buf[grid * 40: grid * 40 + 256].view(64, 4).cpu().tolist()
    print("  CTA 0 warp 2 per tile [wait, strips, tile-end merges, rows merged]:")
    for i in (0, 1, 2, 3, 4, 5, 6, 7, 8, 10, 12, 16, 20, 24, 32, 40, 48, 63):
        print("   tile %2d: %s" % (i, tr[i]))
    # skew of the units that share d-tiles (same group): globaltimer (ns) when the MMA issuer starts tile 8 / 64 / 200
    step = cg
    n_qt = (b + 128 * cg - 1) // (128 * cg)
    for col, tile in ((5, 8), (6, 64), (7, 200)):
        ts = t[0:grid:step, col]
        ts = ts[ts > 0]
        units = len(ts)
        groups = units // n_qt
        if groups == 0:
            continue
        g = ts[: groups * n_qt].view(groups, n_qt)
        spread = (g.max(1).values - g.min(1).values) / 1e3
        print(f"  tile {tile:3d}: spread of start times inside a group: mean {spread.mean().item():.1f} us, max {spread.max().item():.1f} us; "
              f"across all units {(ts.max() - ts.min()).item() / 1e3:.1f} us")
