#!/usr/bin/env python
"""K2p diagnostics: the int8 scan with the full epilogue / TMEM reads only / no epilogue."""
import os, sys
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import torch
import sqe_b200
from sqe_b200 import ops
nat = sqe_b200._native
if os.environ.get("SQE_LIB"):                             # an experimental build of the library (before the first call)
    nat.LIB_PATH = os.path.abspath(os.environ["SQE_LIB"])
rows, b, K = int(sys.argv[1]), int(sys.argv[2]), int(sys.argv[3])
dev = torch.device("cuda", 0)
D = torch.empty((rows, 1024), dtype=torch.bfloat16, device=dev)
gen = torch.Generator(device=dev)
for lo in range(0, rows, 250_000):
    gen.manual_seed(1234 + lo // 250_000)
    ops.normalize_cast(torch.randn((min(250_000, rows - lo), 1024), generator=gen, device=dev), "bf16", out=D[lo:lo + 250_000])
d8, meta = ops.quantize_rows(D)
q = torch.randn((b, 1024), generator=torch.Generator().manual_seed(99)).to(dev)
for mode in [int(m) for m in os.environ.get("K2P_MODES", "0,1,2").split(",")]:
    nat.tuning_set(nat.SQE_TUNE_K2_EPILOGUE_MODE, mode)
    for _ in range(3):
        ops.search_batched_prefiltered(D, d8, meta, q, K)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(5):
        ops.search_batched_prefiltered(D, d8, meta, q, K)
    e1.record()
    torch.cuda.synchronize()
    print(f"lib={os.path.basename(nat.LIB_PATH)} rows={rows} b={b} k={K} epilogue_mode={mode}: {e0.elapsed_time(e1) / 5:.3f} ms per call", flush=True)
nat.tuning_set(nat.SQE_TUNE_K2_EPILOGUE_MODE, 0)
