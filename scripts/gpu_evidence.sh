#!/bin/bash
# round-1 evidence run: full GPU suite, smoke, all bench workloads, ncu launch list + full capture
set -x
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 900 python -m pytest tests -x -q -m gpu 2>&1 | tail -5 > gpurun_out/r1i_tests.log; cat gpurun_out/r1i_tests.log
timeout 300 python __graft_entry__.py smoke 2>&1 | tail -3
timeout 600 python bench.py > gpurun_out/r1i_bench_default.json 2> gpurun_out/r1i_bench_default.err; tail -2 gpurun_out/r1i_bench_default.err; cat gpurun_out/r1i_bench_default.json
timeout 600 python bench.py --workload b1 --no-cpu-baseline > gpurun_out/r1i_bench_b1.json 2>/dev/null; cat gpurun_out/r1i_bench_b1.json
timeout 600 python bench.py --workload b1 --rows 1000000 --dtype fp32 --no-cpu-baseline > gpurun_out/r1i_bench_b1_fp32_1m.json 2>/dev/null; cat gpurun_out/r1i_bench_b1_fp32_1m.json
timeout 600 python bench.py --workload cache64 --no-cpu-baseline --steps 50 > gpurun_out/r1i_bench_cache64.json 2>/dev/null; cat gpurun_out/r1i_bench_cache64.json
for dt in bf16 fp32; do timeout 300 python bench.py --workload ingest --dtype $dt > gpurun_out/r1i_bench_ingest_$dt.json 2>/dev/null; cat gpurun_out/r1i_bench_ingest_$dt.json; done
CMD="python bench.py --rows 2000000 --steps 2 --warmup 1 --no-cpu-baseline --no-e2e"
timeout 300 $CMD > gpurun_out/plain_r1i.log 2>&1 &&
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r1i_launches.csv $CMD > gpurun_out/ncu_r1i_list.log 2>&1
timeout 300 $CMD > gpurun_out/plain_r1i_b.log 2>&1 &&
timeout 900 ncu --set full --clock-control none --import-source on -k regex:topk_batched_kernel -s 1 -c 1 -o gpurun_out/r1i_k2 $CMD > gpurun_out/ncu_r1i_full.log 2>&1
tail -3 gpurun_out/ncu_r1i_full.log
