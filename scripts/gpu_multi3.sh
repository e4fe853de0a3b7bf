#!/bin/bash
set -x
N=${1:-2}
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29544 scripts/multi_check.py 2>&1 | grep -v "OMP_NUM_THREADS\|^\*\*\*" | tail -25 | tee gpurun_out/r1_multi_check_n$N.log
