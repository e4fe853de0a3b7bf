#!/bin/bash
set -x
mkdir -p gpurun_out
timeout 1200 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "cache or graph or ws_session or plugin or cosine" 2>&1 | tail -25 > gpurun_out/r2_cache_tests.log
cat gpurun_out/r2_cache_tests.log
timeout 900 python -m pytest tests/test_reference_handlers_gpu.py -x -q -m gpu -s 2>&1 | tail -12 > gpurun_out/r2_reference_handlers.log
cat gpurun_out/r2_reference_handlers.log
timeout 900 python -m pytest tests/test_gpu_parity_at_size.py -x -q -m gpu -s -k "config5" 2>&1 | tail -8
