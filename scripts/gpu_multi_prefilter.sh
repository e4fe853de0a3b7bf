#!/bin/bash
# N-GPU check of the prefiltered batch-1 path: sharded == single exact shard, default bench (its
# secondary_prefiltered section under torchrun), b=1 bench with --prefilter
set -x
N=${1:-2}
cd "$(dirname "$0")/.."
O=gpurun_out/pfn$N; mkdir -p $O
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1"
SQE_PREFILTER=1 timeout 300 $TR --master-port 29544 scripts/multi_check.py 2>&1 | grep -v "OMP_NUM_THREADS\|^\*\*\*" | tail -8 | tee $O/multi_check_prefilter.log
timeout 400 $TR --master-port 29533 bench.py --gpus $N --no-yardstick > $O/scale_default.json 2> $O/scale_default.err
tail -3 $O/scale_default.err; cat $O/scale_default.json
timeout 300 $TR --master-port 29534 bench.py --gpus $N --workload b1 --prefilter --no-cpu-baseline --steps 200 > $O/scale_b1_prefilter.json 2> $O/scale_b1_prefilter.err
tail -3 $O/scale_b1_prefilter.err; cat $O/scale_b1_prefilter.json
