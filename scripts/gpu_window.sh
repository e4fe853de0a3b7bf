#!/bin/bash
set -x
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
for w in 0 4 8 16; do
for rows in 2000000 10000000; do
CMD="python bench.py --rows $rows --steps 2 --warmup 1 --no-cpu-baseline --no-e2e --no-secondary --k2-window $w"
timeout 300 $CMD > gpurun_out/plain_w.log 2>&1 &&
timeout 600 ncu --metrics dram__bytes_read.sum,gpu__time_duration.sum,sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active --clock-control none -k regex:topk_batched_kernel -s 1 -c 1 --csv --log-file gpurun_out/wx.csv $CMD > /dev/null 2>&1
grep -E "topk_batched" gpurun_out/wx.csv | awk -F'","' '{printf "window '$w' rows '$rows' %s %s | ", $(NF-2), $(NF)}'; echo
done
done
for w in 0 8 16; do
timeout 300 python bench.py --steps 30 --warmup 3 --no-cpu-baseline --no-e2e --no-secondary --k2-window $w 2>/dev/null | python -c "
import sys, json
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('bench10M window $w value', round(d['value']), 'TF', round(d['roofline']['achieved']), d['clocks'])"
done
