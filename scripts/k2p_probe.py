#!/usr/bin/env python
"""K2p probe: time of the three launches and rows scored exactly, vs K2 and K3, for a few shapes.
    python scripts/k2p_probe.py rows b k dtype"""
import os, sys
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import torch
import sqe_b200
from sqe_b200 import ops

rows = int(sys.argv[1]) if len(sys.argv) > 1 else 2_000_000
b = int(sys.argv[2]) if len(sys.argv) > 2 else 1024
K = int(sys.argv[3]) if len(sys.argv) > 3 else 10
DT = sys.argv[4] if len(sys.argv) > 4 else "bf16"
dev = torch.device("cuda", 0)
D = torch.empty((rows, ops.ROW_ELEMS[DT]), dtype=ops.TORCH_DTYPES[DT], device=dev)
gen = torch.Generator(device=dev)
for lo in range(0, rows, 250_000):
    gen.manual_seed(1234 + lo // 250_000)
    x = torch.randn((min(250_000, rows - lo), 1024), generator=gen, device=dev)
    ops.normalize_cast(x, DT, out=D[lo:lo + x.shape[0]])
d8, meta = ops.quantize_rows(D)
q = torch.randn((b, 1024), generator=torch.Generator().manual_seed(99)).to(dev)
qn = ops.normalize_cast(q, DT)
resc = torch.zeros((b,), dtype=torch.int32, device=dev)

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

nat = sqe_b200._native
if os.environ.get("K2P_WINDOW"):
    nat.tuning_set(nat.SQE_TUNE_K2_WINDOW, int(os.environ["K2P_WINDOW"]))
nat.tuning_set(nat.SQE_TUNE_K2_EPILOGUE_MODE, 3)          # report pairs logged
ops.search_batched_prefiltered(D, d8, meta, q, K, rescored=resc)
logged = resc.cpu().numpy().copy()
nat.tuning_set(nat.SQE_TUNE_K2_EPILOGUE_MODE, 0)
t_p = timeit(lambda: ops.search_batched_prefiltered(D, d8, meta, q, K, rescored=resc))
r = resc.cpu().numpy()
line = f"window={os.environ.get('K2P_WINDOW', 'default')} rows={rows} b={b} k={K} {DT}: K2p {t_p:.3f} ms ({b / t_p * 1e3:.0f} q/s), exact rows/query median {int(__import__('numpy').median(r))} max {r.max()} full-scans {(r >= rows).sum()}, logged/query median {int(__import__('numpy').median(logged))} max {logged.max()}"
if DT != "fp32":
    t_2 = timeit(lambda: ops.topk_batched(D, qn, K))
    line += f" | K2 {t_2:.3f} ms ({b / t_2 * 1e3:.0f} q/s) -> x{t_2 / t_p:.2f}"
if b <= 64:
    t_3 = timeit(lambda: ops.topk_gemv(D, qn, K), 3)
    line += f" | K3 {t_3:.3f} ms"
gb = rows * 1040 / 1e9
line += f" | int8 bytes {gb:.2f} GB -> {gb / t_p * 1e3:.0f} GB/s; int8 TOP/s {2 * b * rows * 1024 / t_p / 1e9:.0f}"
print(line, flush=True)
