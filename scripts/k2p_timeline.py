#!/usr/bin/env python
"""K2p tile timeline (EXPERIMENTAL build only: SQE_LIB=.../libsqe_b200_tl.so, a copy of topk_batched_i8.cu in which
epilogue warp 2 of every CTA stores %globaltimer when a tile's accumulator is ready and when its scan ends, into the
buffer of sqe_debug_k2_timers).  Prints where the time of the int8 scan goes along the scan: first tiles, middle,
last tiles, spread of the units' end times -- with the full epilogue and with none (epilogue mode 2)."""
import os, sys
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import numpy as np
import torch
import sqe_b200
from sqe_b200 import ops
nat = sqe_b200._native
nat.LIB_PATH = os.path.abspath(os.environ["SQE_LIB"])
dev = torch.device("cuda", 0)
ROWS = 10_000_000
D = torch.empty((ROWS, 1024), dtype=torch.bfloat16, device=dev)
gen = torch.Generator(device=dev)
for lo in range(0, ROWS, 250_000):
    gen.manual_seed(1234 + lo // 250_000)
    ops.normalize_cast(torch.randn((250_000, 1024), generator=gen, device=dev), "bf16", out=D[lo:lo + 250_000])
d8, meta = ops.quantize_rows(D)
Q = torch.randn((1024, 1024), generator=torch.Generator().manual_seed(99)).to(dev)
buf = torch.zeros((148 * 4096,), dtype=torch.int64, device=dev)
for rows in (10_000_000, 2_500_000, 1_250_000):
    for mode in (0, 2):
        nat.tuning_set(nat.SQE_TUNE_K2_EPILOGUE_MODE, mode)
        for _ in range(3):
            ops.search_batched_prefiltered(D[:rows], d8[:rows], meta[:rows], Q, 10)
        torch.cuda.synchronize()
        buf.zero_()
        nat.load().sqe_debug_k2_timers(buf.data_ptr())
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        ops.search_batched_prefiltered(D[:rows], d8[:rows], meta[:rows], Q, 10)
        e1.record()
        torch.cuda.synchronize()
        nat.load().sqe_debug_k2_timers(None)
        t = buf.cpu().numpy().reshape(148, 4096)
        ctas = [c for c in range(148) if t[c, 0] > 0]
        start = min(t[c, 0] for c in ctas)
        ends = np.array([t[c, 4095] - start for c in ctas]) / 1e3
        n_tiles = [int((t[c, :4094] > 0).sum()) for c in ctas]
        def seg(lo, hi):                                   # mean tile time (us) over tiles [lo, hi) of every CTA
            v = [(t[c, min(hi, n) - 0 - 1] - t[c, lo]) / max(1, min(hi, n) - 1 - lo) for c, n in zip(ctas, n_tiles) if n > lo + 1]
            return float(np.mean(v)) / 1e3 if v else float("nan")
        n = min(n_tiles)
        first_ready = np.array([t[c, 0] - start for c in ctas]) / 1e3
        line = (f"rows={rows} mode={mode}: call {e0.elapsed_time(e1):.3f} ms, {len(ctas)} CTAs x {n}-{max(n_tiles)} tiles; "
                f"first accumulator ready {first_ready.min():.0f}-{first_ready.max():.0f} us; us per tile: "
                f"[0,1) {seg(0, 2):.1f} [1,4) {seg(1, 4):.1f} [4,16) {seg(4, 16):.1f} [16,64) {seg(16, 64):.1f} "
                f"[64,n/2) {seg(64, n // 2):.2f} [n/2,n-16) {seg(n // 2, n - 16):.2f} last16 {seg(n - 16, n):.2f}; "
                f"scan ends (us after the first stamp) min {ends.min():.0f} median {np.median(ends):.0f} max {ends.max():.0f}")
        print(line, flush=True)
nat.tuning_set(nat.SQE_TUNE_K2_EPILOGUE_MODE, 0)
