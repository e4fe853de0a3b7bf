#!/bin/bash
set -x
mkdir -p gpurun_out
python scripts/k2p_probe.py 1000000 1024 10 bf16 > gpurun_out/r2_k2p_ncu_plain.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:topk_batched_i8 -s 3 -c 1 -f -o gpurun_out/r2_k2p_i8 python scripts/k2p_probe.py 1000000 1024 10 bf16 > gpurun_out/r2_k2p_ncu.log 2>&1
tail -3 gpurun_out/r2_k2p_ncu_plain.log; tail -3 gpurun_out/r2_k2p_ncu.log; ls -la gpurun_out/*.ncu-rep
