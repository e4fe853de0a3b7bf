#!/usr/bin/env python
"""Interleaved A/B of the tensor-core kernel variants on ONE box (power-capped clocks differ
between boxes), with a cuBLAS GEMM of the same shape (no top-k, output written to HBM) as the
yardstick of what the power cap allows.   python scripts/k2_ab.py [rows] [b] [iters]"""
import os, subprocess, sys, threading, time
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import torch
import sqe_b200
from sqe_b200 import ops
nat = sqe_b200._native

rows = int(sys.argv[1]) if len(sys.argv) > 1 else 10_000_000
b = int(sys.argv[2]) if len(sys.argv) > 2 else 1024
iters = int(sys.argv[3]) if len(sys.argv) > 3 else 40
dev = torch.device("cuda", 0)
print(subprocess.run(["nvidia-smi", "--query-gpu=power.limit,power.default_limit,power.max_limit,clocks.max.sm,temperature.gpu",
                      "--format=csv"], capture_output=True, text=True).stdout)
D = torch.empty((rows, 1024), dtype=torch.bfloat16, device=dev)
gen = torch.Generator(device=dev)
for lo in range(0, rows, 250_000):
    gen.manual_seed(lo)
    x = torch.randn((min(250_000, rows - lo), 1024), generator=gen, device=dev)
    ops.normalize_cast(x, "bf16", out=D[lo:lo + x.shape[0]])
Q = ops.normalize_cast(torch.randn((b, 1024), generator=gen, device=dev), "bf16")
flops = 2.0 * rows * b * 1024

samples = []
stop = False
def sampler():
    p = subprocess.Popen(["nvidia-smi", "--query-gpu=clocks.sm,power.draw", "--format=csv,noheader,nounits", "-lms", "20"],
                         stdout=subprocess.PIPE, text=True)
    for line in p.stdout:
        samples.append((time.time(), line.strip()))
        if stop:
            break
    p.terminate()
threading.Thread(target=sampler, daemon=True).start()
time.sleep(1.5)

def clocks(t0, t1):
    sm, pw = [], []
    for ts, ln in samples:
        if t0 <= ts <= t1:
            a = ln.split(",")
            sm.append(float(a[0])); pw.append(float(a[1]))
    if not sm:
        return "no samples"
    sm.sort(); pw.sort()
    return f"sm {sm[len(sm)//2]:.0f} MHz (min {sm[0]:.0f}) power {pw[len(pw)//2]:.0f} W n={len(sm)}"

def run(name, fn):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t0 = time.time()
    e0.record()
    for _ in range(iters):
        fn()
    e1.record()
    torch.cuda.synchronize()
    t1 = time.time()
    ms = e0.elapsed_time(e1) / iters
    print(f"{name:28s} {ms:8.3f} ms  {flops / (ms * 1e-3) / 1e12:7.1f} TFLOP/s  {b / (ms * 1e-3):9.0f} q/s   {clocks(t0, t1)}", flush=True)

def k2(cg):
    def f():
        nat.tuning_set(nat.SQE_TUNE_K2_CTA_GROUP, cg)
        ops.topk_batched(D, Q, 10)
    return f

chunk = 1_000_000
out = torch.empty((b, chunk), dtype=torch.bfloat16, device=dev)
def cublas():
    for lo in range(0, rows, chunk):
        torch.matmul(Q, D[lo:lo + chunk].T, out=out[:, : min(chunk, rows - lo)])

for rep in range(2):
    run("K2 cta_group=2", k2(2))
    run("K2 cta_group=1", k2(1))
    run("cuBLAS GEMM only (bf16 out)", cublas)
for mode in (2, 1):
    nat.tuning_set(nat.SQE_TUNE_K2_EPILOGUE_MODE, mode)
    run(f"K2 cg2 diag epilogue_mode={mode}", k2(2))
nat.tuning_set(nat.SQE_TUNE_K2_EPILOGUE_MODE, 0)
stop = True
