#!/bin/bash
# launch list + ncu --set full of the prefiltered batch-1 path (K3p) at the benchmarked size
set -x
cd "$(dirname "$0")/.."
O=gpurun_out/ncupf; mkdir -p $O
CMD="python bench.py --workload b1 --prefilter --steps 3 --warmup 1 --no-cpu-baseline --no-e2e"
timeout 300 $CMD > $O/plain.log 2>&1 || exit 1
tail -c 600 $O/plain.log
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:"coarse_scan|rescore|normalize_cast|quantize_rows|topk_gemv" -c 400 --csv --log-file $O/launches.csv $CMD > $O/ncu_list.log 2>&1
tail -2 $O/ncu_list.log
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"coarse_scan_kernel|rescore_kernel" -s 4 -c 2 -o $O/pf_10m $CMD > $O/ncu_full.log 2>&1
tail -2 $O/ncu_full.log
ls -la $O
