#!/bin/bash
set -x
N=${1:-8}
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29544 scripts/multi_check.py 2>&1 | grep -v "OMP_NUM_THREADS\|^\*\*\*" | tail -12 | tee gpurun_out/r1_multi_check_n$N.log
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus $N > gpurun_out/r1_scale3_default_n$N.json 2> gpurun_out/r1_scale3_default_n$N.err
tail -3 gpurun_out/r1_scale3_default_n$N.err; cat gpurun_out/r1_scale3_default_n$N.json
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29534 bench.py --gpus $N --workload b1 --no-cpu-baseline > gpurun_out/r1_scale3_b1_n$N.json 2> gpurun_out/r1_scale3_b1_n$N.err
tail -3 gpurun_out/r1_scale3_b1_n$N.err; cat gpurun_out/r1_scale3_b1_n$N.json
