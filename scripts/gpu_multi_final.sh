#!/bin/bash
# multi-GPU evidence: sharded-search check, default bench, b=1 bench, BASELINE configs[3] (100M fp16, b=256, k=100)
set -x
N=${1:-8}
WHAT=${2:-all}
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1"
if [ "$WHAT" = "all" ]; then
timeout 600 $TR --master-port 29544 scripts/multi_check.py 2>&1 | grep -v "OMP_NUM_THREADS\|^\*\*\*" | tail -8 | tee gpurun_out/r1_multi_check_n$N.log
timeout 900 $TR --master-port 29533 bench.py --gpus $N > gpurun_out/r1_final_scale_default_n$N.json 2> gpurun_out/r1_final_scale_default_n$N.err
tail -3 gpurun_out/r1_final_scale_default_n$N.err; cat gpurun_out/r1_final_scale_default_n$N.json
timeout 900 $TR --master-port 29534 bench.py --gpus $N --workload b1 --no-cpu-baseline --steps 200 > gpurun_out/r1_final_scale_b1_n$N.json 2> gpurun_out/r1_final_scale_b1_n$N.err
tail -3 gpurun_out/r1_final_scale_b1_n$N.err; cat gpurun_out/r1_final_scale_b1_n$N.json
fi
timeout 1200 $TR --master-port 29535 bench.py --gpus $N --workload b1024 --batch 256 --k 100 --dtype fp16 --rows 100000000 --no-cpu-baseline --no-secondary > gpurun_out/r1_final_cfg4_n$N.json 2> gpurun_out/r1_final_cfg4_n$N.err
tail -3 gpurun_out/r1_final_cfg4_n$N.err; cat gpurun_out/r1_final_cfg4_n$N.json
