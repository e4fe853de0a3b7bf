#!/bin/bash
# Round 2 final, N GPUs: the default bench line as the driver launches it (+ the reference arm on rank 0).
set -x
mkdir -p gpurun_out
N=${1:-8}
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29561 \
    bench.py --gpus $N --steps 20 --warmup 3 > gpurun_out/r2_scale_default_n${N}_k2p.json 2> gpurun_out/r2_scale_default_n${N}_k2p.err
tail -c 3500 gpurun_out/r2_scale_default_n${N}_k2p.json; tail -3 gpurun_out/r2_scale_default_n${N}_k2p.err
