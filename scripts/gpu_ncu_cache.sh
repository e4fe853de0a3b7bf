#!/bin/bash
set -x
cd "$(dirname "$0")/.."
O=gpurun_out/evidence3; mkdir -p $O
CMD="python bench.py --workload cache64 --steps 3 --warmup 1 --no-cpu-baseline --no-e2e"
timeout 300 $CMD > $O/plain.log 2>&1 &&
timeout 600 ncu --set full --clock-control none --import-source on -k regex:topk_batched_kernel -s 2 -c 1 -o $O/cache64 $CMD > $O/ncu.log 2>&1
tail -2 $O/ncu.log
