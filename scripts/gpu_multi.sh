#!/bin/bash
# multi-GPU bench: N ranks on one box (torchrun, NCCL)
set -x
N=${1:-2}
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
nvidia-smi topo -m | head -12
for wl in b1024 b1; do
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus $N --workload $wl --no-cpu-baseline > gpurun_out/r1_scale_${wl}_n$N.json 2> gpurun_out/r1_scale_${wl}_n$N.err
tail -5 gpurun_out/r1_scale_${wl}_n$N.err; cat gpurun_out/r1_scale_${wl}_n$N.json
done
