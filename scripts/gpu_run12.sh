#!/bin/bash
set -x
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 900 python -m pytest tests -x -q -m gpu 2>&1 | tail -6
timeout 300 python bench.py --workload config1 > gpurun_out/r1j_bench_config1.json 2> gpurun_out/r1j_bench_config1.err; tail -3 gpurun_out/r1j_bench_config1.err; cat gpurun_out/r1j_bench_config1.json
timeout 300 python bench.py --workload cache64 --no-cpu-baseline --steps 50 > gpurun_out/r1j_bench_cache64.json 2>/dev/null; cat gpurun_out/r1j_bench_cache64.json
