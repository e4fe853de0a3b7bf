#!/bin/bash
set -x
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "batched or tensor_path or cache" 2>&1 | tail -8
timeout 600 python bench.py --workload b1024 --batch 256 --k 100 --dtype fp16 --rows 12500000 --no-cpu-baseline --no-secondary > gpurun_out/r1g_bench_cfg4_shard.json 2> gpurun_out/r1g_bench_cfg4_shard.err
tail -3 gpurun_out/r1g_bench_cfg4_shard.err; cat gpurun_out/r1g_bench_cfg4_shard.json
timeout 600 python bench.py --workload b1024 --rows 1250000 --no-cpu-baseline --no-secondary --steps 50 > gpurun_out/r1g_bench_b1024_1p25m.json 2> gpurun_out/r1g_bench_b1024_1p25m.err
tail -3 gpurun_out/r1g_bench_b1024_1p25m.err; cat gpurun_out/r1g_bench_b1024_1p25m.json
