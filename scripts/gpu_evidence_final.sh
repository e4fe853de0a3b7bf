#!/bin/bash
# final single-GPU evidence run of the round (under gpurun): full GPU suite, smoke, every bench
# workload, the reference arm, ncu launch list of the exact default command, --set full captures
set -x
cd "$(dirname "$0")/.."
O=gpurun_out/final
mkdir -p $O
timeout 1200 python -m pytest tests -x -q -m gpu 2>&1 | tail -5 > $O/tests.log; cat $O/tests.log
timeout 300 python __graft_entry__.py smoke 2>&1 | tail -2 | tee $O/smoke.log
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > $O/bench_reference.json 2>/dev/null; cat $O/bench_reference.json
timeout 600 python bench.py > $O/bench_default.json 2> $O/bench_default.err; tail -2 $O/bench_default.err; cat $O/bench_default.json
timeout 600 python bench.py --workload b1 > $O/bench_b1.json 2>/dev/null; cat $O/bench_b1.json
timeout 600 python bench.py --workload b1 --rows 1000000 --dtype fp32 --no-cpu-baseline > $O/bench_b1_fp32_1m.json 2>/dev/null; cat $O/bench_b1_fp32_1m.json
timeout 600 python bench.py --workload cache64 --steps 50 > $O/bench_cache64.json 2>/dev/null; cat $O/bench_cache64.json
timeout 600 python bench.py --workload cache64 --dtype bf16x2 --steps 50 --no-cpu-baseline > $O/bench_cache64_bf16x2.json 2>/dev/null; cat $O/bench_cache64_bf16x2.json
for dt in bf16 fp32 bf16x2; do timeout 300 python bench.py --workload ingest --dtype $dt > $O/bench_ingest_$dt.json 2>/dev/null; cat $O/bench_ingest_$dt.json; done
timeout 300 python bench.py --workload config1 > $O/bench_config1.json 2>/dev/null; cat $O/bench_config1.json
timeout 600 python bench.py --workload serve > $O/bench_serve.json 2>/dev/null; cat $O/bench_serve.json
timeout 300 python bench.py --no-cpu-baseline --batch 256 --k 100 --dtype fp16 --rows 12500000 --no-secondary > $O/bench_cfg4_shard_1gpu.json 2>/dev/null; cat $O/bench_cfg4_shard_1gpu.json
CMD="python bench.py --no-cpu-baseline"
timeout 600 $CMD > $O/plain_default.json 2> $O/plain_default.err &&
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file $O/launches_default.csv $CMD > $O/ncu_list.log 2>&1
CMD2="python bench.py --rows 2000000 --steps 2 --warmup 1 --no-cpu-baseline --no-e2e --no-yardstick"
timeout 300 $CMD2 > $O/plain_2m.log 2>&1 &&
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"topk_batched_kernel|topk_gemv_kernel" -s 1 -c 3 -o $O/k2_gemv $CMD2 > $O/ncu_full.log 2>&1
tail -2 $O/ncu_full.log
ls -la $O
