#!/bin/bash
# ncu --set full capture of the tensor-core kernel on a 2M-row shard (b=1024, CTA pairs)
set -x
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
CMD="python bench.py --workload b1024 --rows 2000000 --steps 2 --warmup 1 --no-cpu-baseline --no-e2e --k2-cta-group ${1:-2}"
timeout 300 $CMD > gpurun_out/plain_k2c.log 2>&1 &&
timeout 900 ncu --set full --clock-control none --import-source on -k regex:topk_batched_kernel -s 1 -c 1 -o gpurun_out/${2:-r1c_k2} $CMD > gpurun_out/ncu_k2c_full.log 2>&1
tail -3 gpurun_out/ncu_k2c_full.log
