#!/bin/bash
set -x
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "batched or tensor_path or cache" 2>&1 | tail -8
for i in 1 2; do
timeout 600 python bench.py --workload cache64 --no-cpu-baseline --steps 50 > gpurun_out/r1f_bench_cache64.json 2> gpurun_out/r1f_bench_cache64.err
tail -3 gpurun_out/r1f_bench_cache64.err; cat gpurun_out/r1f_bench_cache64.json
done
