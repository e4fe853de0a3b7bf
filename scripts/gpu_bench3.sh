#!/bin/bash
set -x
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
for cg in 2 1; do
timeout 600 python bench.py --workload b1024 --no-cpu-baseline --k2-cta-group $cg > gpurun_out/r1d_bench_b1024_cg$cg.json 2> gpurun_out/r1d_bench_b1024_cg$cg.err
tail -3 gpurun_out/r1d_bench_b1024_cg$cg.err; cat gpurun_out/r1d_bench_b1024_cg$cg.json
done
timeout 600 python bench.py --workload cache64 --no-cpu-baseline > gpurun_out/r1d_bench_cache64.json 2> gpurun_out/r1d_bench_cache64.err
tail -3 gpurun_out/r1d_bench_cache64.err; cat gpurun_out/r1d_bench_cache64.json
