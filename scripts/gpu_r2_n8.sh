#!/bin/bash
# Round 2, N GPUs: sharded parity vs the oracle, the one-query A/B, the default bench line.
set -x
mkdir -p gpurun_out
N=${1:-8}
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29541 \
    tests/multi_gpu_worker.py --exchange p2p 2>&1 | grep -v "^\*\*\*\|OMP_NUM" | tail -12 > gpurun_out/r2_multi_parity_n$N.log
cat gpurun_out/r2_multi_parity_n$N.log
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29542 \
    scripts/multi_b1_ab.py 2> gpurun_out/r2_b1_ab_n$N.err | tail -1 > gpurun_out/r2_b1_ab_n$N.json
cat gpurun_out/r2_b1_ab_n$N.json
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29543 \
    bench.py --gpus $N --steps 20 --warmup 3 > gpurun_out/r2_scale_default_n$N.json 2> gpurun_out/r2_scale_default_n$N.err
tail -c 2500 gpurun_out/r2_scale_default_n$N.json; tail -3 gpurun_out/r2_scale_default_n$N.err
