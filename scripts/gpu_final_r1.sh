#!/bin/bash
# last evidence run of round 1: full GPU suite, the default bench, its ncu launch list, b=1 --prefilter
set -x
cd "$(dirname "$0")/.."
O=gpurun_out/final_r1; mkdir -p $O
timeout 400 python -m pytest tests -x -q -m gpu 2>&1 | tail -5 > $O/tests.log; cat $O/tests.log
CMD="python bench.py --no-cpu-baseline"
timeout 300 $CMD > $O/plain_default.json 2> $O/plain_default.err &&
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file $O/launches_default.csv $CMD > $O/ncu_list.log 2>&1
tail -2 $O/ncu_list.log; cat $O/plain_default.json
timeout 200 python bench.py --workload b1 --prefilter --no-cpu-baseline > $O/bench_b1_prefilter.json 2>/dev/null; cat $O/bench_b1_prefilter.json
