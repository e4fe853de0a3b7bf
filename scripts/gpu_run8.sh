#!/bin/bash
set -x
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 900 python -m pytest tests -x -q -m gpu 2>&1 | tail -8
timeout 600 python bench.py --workload b1 --no-cpu-baseline --rows 1250000 --steps 200 > gpurun_out/r1h_bench_b1_1p25m.json 2> gpurun_out/r1h_bench_b1_1p25m.err
tail -3 gpurun_out/r1h_bench_b1_1p25m.err; cat gpurun_out/r1h_bench_b1_1p25m.json
