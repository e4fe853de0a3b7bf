#!/bin/bash
# Round 2, 2 GPUs: sharded parity vs the oracle (fused + kernel exchange, NCCL fallback), default bench line.
set -x
mkdir -p gpurun_out
N=${1:-2}
timeout 900 python -m pytest tests/test_gpu_parity_at_size.py -x -q -m gpu -s -k two_rank 2>&1 | tail -25 > gpurun_out/r2_multi_parity_n$N.log
cat gpurun_out/r2_multi_parity_n$N.log
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29533 \
    bench.py --gpus $N --steps 20 --warmup 3 > gpurun_out/r2_scale_default_n$N.json 2> gpurun_out/r2_scale_default_n$N.err
tail -c 3000 gpurun_out/r2_scale_default_n$N.json; tail -5 gpurun_out/r2_scale_default_n$N.err
