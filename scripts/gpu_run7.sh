#!/bin/bash
set -x
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "batched or tensor_path or cache" 2>&1 | tail -8
timeout 300 python scripts/k2_timers.py 2500000 256 2 100 fp16 2>&1 | grep -E "epilogue_mode|mma_|epi_warp2_flush|epi_warp2_total|tile  [0-8]"
timeout 300 python scripts/k2_timers.py 2000000 1024 2 2>&1 | grep -E "epilogue_mode|mma_|epi_warp2_flush|epi_warp2_total|tile  [0-4]"
