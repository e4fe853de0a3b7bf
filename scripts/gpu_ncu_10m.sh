#!/bin/bash
# ncu --set full of the two headline kernels at the benchmarked size (10M x 1024 bf16)
set -x
cd "$(dirname "$0")/.."
O=gpurun_out/ncu10m; mkdir -p $O
CMD="python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-e2e --no-yardstick --no-secondary"
timeout 300 $CMD > $O/plain_k2.log 2>&1 &&
timeout 600 ncu --set full --clock-control none --import-source on -k regex:topk_batched_kernel -s 2 -c 1 -o $O/k2_10m $CMD > $O/ncu_k2.log 2>&1
tail -2 $O/ncu_k2.log
CMD2="python bench.py --workload b1 --steps 3 --warmup 1 --no-cpu-baseline --no-e2e"
timeout 300 $CMD2 > $O/plain_b1.log 2>&1 &&
timeout 600 ncu --set full --clock-control none --import-source on -k regex:topk_gemv_kernel -s 2 -c 1 -o $O/gemv_10m $CMD2 > $O/ncu_gemv.log 2>&1
tail -2 $O/ncu_gemv.log
ls -la $O
