#!/bin/bash
set -x
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -x -q -m gpu -s 2>&1 | tail -25 > gpurun_out/r2_k2p_full_tests.log
cat gpurun_out/r2_k2p_full_tests.log
