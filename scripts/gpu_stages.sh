#!/bin/bash
set -x
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
cp semantic-query-engine_b200/lib/libsqe_b200.so /tmp/lib_st5.so
for st in 5 4 3; do
if [ $st != 5 ]; then cp semantic-query-engine_b200/lib/libsqe_b200_st$st.so.bak semantic-query-engine_b200/lib/libsqe_b200.so; else cp /tmp/lib_st5.so semantic-query-engine_b200/lib/libsqe_b200.so; fi
touch semantic-query-engine_b200/lib/libsqe_b200.so
for rows in 2000000 10000000; do
CMD="python bench.py --rows $rows --steps 2 --warmup 1 --no-cpu-baseline --no-e2e --no-secondary"
timeout 300 $CMD > gpurun_out/plain_s.log 2>&1 &&
timeout 600 ncu --metrics dram__bytes_read.sum,gpu__time_duration.sum,sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active --clock-control none -k regex:topk_batched_kernel -s 1 -c 1 --csv --log-file gpurun_out/sx.csv $CMD > /dev/null 2>&1
grep -E "topk_batched" gpurun_out/sx.csv | awk -F'","' '{printf "stages '$st' rows '$rows' %s %s | ", $(NF-2), $(NF)}'; echo
done
timeout 300 python bench.py --steps 30 --warmup 3 --no-cpu-baseline --no-e2e --no-secondary 2>/dev/null | python -c "
import sys, json
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('bench10M stages $st value', round(d['value']), 'TF', round(d['roofline']['achieved']), d['clocks'])"
done
cp /tmp/lib_st5.so semantic-query-engine_b200/lib/libsqe_b200.so
