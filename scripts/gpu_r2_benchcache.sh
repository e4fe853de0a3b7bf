#!/bin/bash
set -x
mkdir -p gpurun_out
for dt in bf16 fp32; do
  timeout 600 python bench.py --workload cache64 --dtype $dt --no-traffic-probe > gpurun_out/r2_bench_cache64_$dt.json 2> gpurun_out/r2_bench_cache64_$dt.err
  tail -c 2500 gpurun_out/r2_bench_cache64_$dt.json; tail -3 gpurun_out/r2_bench_cache64_$dt.err
done
timeout 600 python bench.py --workload cachemut > gpurun_out/r2_bench_cachemut.json 2> gpurun_out/r2_bench_cachemut.err
cat gpurun_out/r2_bench_cachemut.json; tail -3 gpurun_out/r2_bench_cachemut.err
timeout 600 python bench.py --workload serve > gpurun_out/r2_bench_serve.json 2> gpurun_out/r2_bench_serve.err
tail -c 1500 gpurun_out/r2_bench_serve.json; tail -3 gpurun_out/r2_bench_serve.err
