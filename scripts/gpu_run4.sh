#!/bin/bash
set -x
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 900 python -m pytest tests -x -q -m gpu 2>&1 | tail -15 > gpurun_out/r1e_tests.log
cat gpurun_out/r1e_tests.log
timeout 600 python bench.py --workload cache64 --no-cpu-baseline > gpurun_out/r1e_bench_cache64.json 2> gpurun_out/r1e_bench_cache64.err
tail -3 gpurun_out/r1e_bench_cache64.err; cat gpurun_out/r1e_bench_cache64.json
timeout 600 python bench.py > gpurun_out/r1e_bench_default.json 2> gpurun_out/r1e_bench_default.err
tail -3 gpurun_out/r1e_bench_default.err; cat gpurun_out/r1e_bench_default.json
timeout 600 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/r1e_bench_reference.json 2> gpurun_out/r1e_bench_reference.err
cat gpurun_out/r1e_bench_reference.json
