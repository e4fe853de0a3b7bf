#!/bin/bash
set -x
mkdir -p gpurun_out
timeout 1200 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -s -k "batched_prefiltered or index_with_prefilter" 2>&1 | tail -6 > gpurun_out/r2_k2p_tests.log
cat gpurun_out/r2_k2p_tests.log
( timeout 300 python scripts/k2p_modes.py 2500000 1024 10 ) 2>&1 | grep "rows=" | tee gpurun_out/r2_k2p_modes.txt
( for args in "1000000 64 1 bf16" "2500000 1024 10 bf16" "10000000 1024 10 bf16" "12500000 256 100 fp16" "10000000 128 10 bf16" "10000000 256 10 bf16"; do
    timeout 300 python scripts/k2p_probe.py $args 2>&1 | tail -1
  done ) | tee gpurun_out/r2_k2p_probe.txt
