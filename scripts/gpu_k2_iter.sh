#!/bin/bash
# K2 iteration loop: parity tests of both kernel forms, then role timers
set -x
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "batched or tensor_path or full_size" 2>&1 | tail -30 > gpurun_out/k2i_tests.log
cat gpurun_out/k2i_tests.log
timeout 300 python scripts/k2_timers.py 2000000 1024 2>&1 | grep -E "epilogue_mode|mma_wait|epi_|prod" > gpurun_out/k2i_timers_2m.log
cat gpurun_out/k2i_timers_2m.log
