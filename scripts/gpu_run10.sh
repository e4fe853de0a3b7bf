#!/bin/bash
set -x
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "batched or tensor_path or cache or full_size" 2>&1 | tail -5
timeout 300 python scripts/k2_timers.py 2000000 1024 0 2>&1 | grep -E "epilogue_mode=0 cta|spread|mma_wait|prod_wait"
for cg in 2 1; do
CMD="python bench.py --rows 2000000 --steps 2 --warmup 1 --no-cpu-baseline --no-e2e --no-secondary --k2-cta-group $cg"
timeout 300 $CMD > gpurun_out/plain_hint.log 2>&1 &&
timeout 600 ncu --metrics dram__bytes_read.sum,lts__t_sector_op_read_hit_rate.pct,gpu__time_duration.sum,sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active --clock-control none -k regex:topk_batched_kernel -s 1 -c 1 --csv --log-file gpurun_out/hintx.csv $CMD > /dev/null 2>&1
grep -E "topk_batched" gpurun_out/hintx.csv | awk -F'","' '{print "cg '$cg'", $(NF-2), $(NF)}'
done
timeout 600 python scripts/k2_ab.py 10000000 1024 30 2>&1 | grep -E "K2|cuBLAS"
