#!/bin/bash
# Round 2: fused scan + exchange (b <= 2), K3p raw-query form.  1 GPU part.
set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "fused or gemv or prefilter or graph or exchange or index or cache" 2>&1 | tail -15 > gpurun_out/r2_fused_tests.log
cat gpurun_out/r2_fused_tests.log
B="python bench.py --no-cpu-baseline --no-cfg4 --no-traffic-probe --no-yardstick"
timeout 600 $B --workload b1 --steps 50 > gpurun_out/r2_fused_b1.json 2> gpurun_out/r2_fused_b1.err; tail -c 1500 gpurun_out/r2_fused_b1.json
timeout 600 $B --workload b1 --prefilter --steps 50 > gpurun_out/r2_fused_b1_pf.json 2> gpurun_out/r2_fused_b1_pf.err; tail -c 1500 gpurun_out/r2_fused_b1_pf.json
