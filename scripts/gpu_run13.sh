#!/bin/bash
set -x
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 900 python -m pytest tests -x -q -m gpu 2>&1 | tail -5
timeout 600 python bench.py > gpurun_out/r1k_bench_default.json 2> gpurun_out/r1k_bench_default.err; tail -2 gpurun_out/r1k_bench_default.err; cat gpurun_out/r1k_bench_default.json
timeout 600 python bench.py --workload b1024 --batch 256 --k 100 --dtype fp16 --rows 12500000 --no-cpu-baseline --no-secondary > gpurun_out/r1k_bench_cfg4_shard.json 2>/dev/null; cat gpurun_out/r1k_bench_cfg4_shard.json
timeout 600 python bench.py --workload b1024 --k 100 --rows 2500000 --no-cpu-baseline --no-secondary > gpurun_out/r1k_bench_k100.json 2>/dev/null; cat gpurun_out/r1k_bench_k100.json
timeout 600 python scripts/k2_ab.py 10000000 1024 30 2>&1 | grep -E "K2 cta_group=2|cuBLAS"
CMD="python bench.py --rows 2000000 --steps 2 --warmup 1 --no-cpu-baseline --no-e2e"
timeout 300 $CMD > gpurun_out/plain_r1k.log 2>&1 &&
timeout 900 ncu --set full --clock-control none --import-source on -k regex:topk_batched_kernel -s 1 -c 1 -o gpurun_out/r1k_k2 $CMD > gpurun_out/ncu_r1k_full.log 2>&1
tail -2 gpurun_out/ncu_r1k_full.log
