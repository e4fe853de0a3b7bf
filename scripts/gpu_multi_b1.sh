#!/bin/bash
set -x
N=${1:-8}
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1"
timeout 600 $TR --master-port 29544 scripts/multi_check.py 2>&1 | grep -v "OMP_NUM_THREADS\|^\*\*\*" | tail -6 | tee gpurun_out/r1_multi_check_n$N.log
timeout 900 $TR --master-port 29534 bench.py --gpus $N --workload b1 --no-cpu-baseline --steps 300 > gpurun_out/r1_final_scale_b1_n$N.json 2> gpurun_out/r1_final_scale_b1_n$N.err
tail -3 gpurun_out/r1_final_scale_b1_n$N.err; cat gpurun_out/r1_final_scale_b1_n$N.json
