#!/bin/bash
set -x
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
for cfg in "4 2" "0 1" "4 1"; do
set -- $cfg
CMD="python bench.py --rows 2000000 --steps 2 --warmup 1 --no-cpu-baseline --no-e2e --no-secondary --k2-d-hint $1 --k2-cta-group $2"
timeout 300 $CMD > gpurun_out/plain_hint.log 2>&1 &&
timeout 600 ncu --metrics dram__bytes_read.sum,lts__t_sector_op_read_hit_rate.pct,gpu__time_duration.sum,lts__t_sectors_srcunit_ltcfabric_lookup_miss.sum,lts__t_sectors_srcunit_ltcfabric_lookup_hit.sum --clock-control none -k regex:topk_batched_kernel -s 1 -c 1 --csv --log-file gpurun_out/hintx.csv $CMD > /dev/null 2>&1
grep -E "topk_batched" gpurun_out/hintx.csv | awk -F'","' '{print "hint '$1' cg '$2'", $(NF-2), $(NF)}'
done
