#!/bin/bash
# full GPU suite + headline benches + ncu evidence for the tensor-core kernel
set -x
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 900 python -m pytest tests -x -q -m gpu 2>&1 | tail -30 > gpurun_out/r1b_tests.log
cat gpurun_out/r1b_tests.log
timeout 600 python bench.py --workload b1024 > gpurun_out/r1b_bench_b1024.json 2> gpurun_out/r1b_bench_b1024.err
tail -3 gpurun_out/r1b_bench_b1024.err; cat gpurun_out/r1b_bench_b1024.json
timeout 600 python bench.py --workload b1 --no-cpu-baseline > gpurun_out/r1b_bench_b1.json 2> gpurun_out/r1b_bench_b1.err
cat gpurun_out/r1b_bench_b1.json
timeout 600 python bench.py --workload cache64 --no-cpu-baseline > gpurun_out/r1b_bench_cache64.json 2> gpurun_out/r1b_bench_cache64.err
tail -3 gpurun_out/r1b_bench_cache64.err; cat gpurun_out/r1b_bench_cache64.json
CMD="python bench.py --workload b1024 --rows 2000000 --steps 2 --warmup 1 --no-cpu-baseline --no-e2e"
timeout 300 $CMD > gpurun_out/plain_k2.log 2>&1 &&
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r1b_launches_b1024.csv $CMD > gpurun_out/ncu_k2_list.log 2>&1
timeout 300 $CMD > gpurun_out/plain_k2b.log 2>&1 &&
timeout 900 ncu --set full --clock-control none --import-source on -k regex:topk_batched_kernel -s 1 -c 2 -o gpurun_out/r1b_k2 $CMD > gpurun_out/ncu_k2_full.log 2>&1
tail -5 gpurun_out/ncu_k2_full.log
