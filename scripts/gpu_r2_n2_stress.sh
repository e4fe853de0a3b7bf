#!/bin/bash
# Round 2, 2 GPUs: 10,000 skewed exchange epochs + parity vs the oracle.
mkdir -p gpurun_out
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29551 \
    tests/multi_gpu_worker.py --exchange p2p --stress ${1:-10000} > gpurun_out/r2_exchange_stress_full.log 2>&1
grep -v "^\*\*\*\|OMP_NUM\|NCCL version" gpurun_out/r2_exchange_stress_full.log | grep -B2 -A12 "Error\|rank ok" | cut -c1-1500 | head -80 > gpurun_out/r2_exchange_stress_n2.log
cat gpurun_out/r2_exchange_stress_n2.log
