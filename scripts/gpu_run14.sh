#!/bin/bash
set -x
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 900 python -m pytest tests -x -q -m gpu 2>&1 | tail -5
timeout 600 python bench.py --workload serve > gpurun_out/r1l_bench_serve.json 2> gpurun_out/r1l_bench_serve.err; tail -3 gpurun_out/r1l_bench_serve.err; cat gpurun_out/r1l_bench_serve.json
timeout 600 python bench.py --workload b1 > gpurun_out/r1l_bench_b1.json 2>/dev/null; cat gpurun_out/r1l_bench_b1.json
