#!/bin/bash
mkdir -p gpurun_out
( for args in "2500000 1024 10 bf16" "10000000 1024 10 bf16" "10000000 128 10 bf16" "1000000 64 1 bf16"; do
    timeout 300 python scripts/k2p_probe.py $args 2>&1 | tail -1
  done ) | tee gpurun_out/r2_k2p_probe_quick.txt
