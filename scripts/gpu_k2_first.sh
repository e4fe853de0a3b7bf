#!/bin/bash
# first contact of the tcgen05 kernel with a B200: small cases first, each under its own timeout
set -x
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,memory.total,clocks.max.sm --format=csv
timeout 300 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "batched or tensor_path" 2>&1 | tail -40 > gpurun_out/k2_tests.log
cat gpurun_out/k2_tests.log
timeout 300 python bench.py --workload b1024 --rows 2000000 --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/k2_bench_2m.json 2> gpurun_out/k2_bench_2m.err
tail -5 gpurun_out/k2_bench_2m.err; cat gpurun_out/k2_bench_2m.json
