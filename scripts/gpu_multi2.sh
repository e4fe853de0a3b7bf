#!/bin/bash
set -x
N=${1:-2}
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "exchange" 2>&1 | tail -15
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29544 scripts/multi_check.py 2>&1 | grep -v "OMP_NUM_THREADS\|^\*\*\*" | tail -25 | tee gpurun_out/r1_multi_check_n$N.log
for wl in b1 b1024; do
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus $N --workload $wl --no-cpu-baseline > gpurun_out/r1_scale2_${wl}_n$N.json 2> gpurun_out/r1_scale2_${wl}_n$N.err
tail -3 gpurun_out/r1_scale2_${wl}_n$N.err; cat gpurun_out/r1_scale2_${wl}_n$N.json
done
