#!/bin/bash
mkdir -p gpurun_out
( for rows in 500000 1000000 2000000 4000000; do for b in 32 64 128 256 512 1024; do
    timeout 120 python scripts/k2p_probe.py $rows $b 10 bf16 2>&1 | tail -1 | sed -E 's/exact rows.*\| K2 /| K2 /; s/\| int8 bytes.*//'
  done; done
  for rows in 1000000 4000000; do for b in 64 256; do
    timeout 120 python scripts/k2p_probe.py $rows $b 1 bf16 2>&1 | tail -1 | sed -E 's/exact rows.*\| K2 /| K2 /; s/\| int8 bytes.*//'
  done; done ) | tee gpurun_out/r2_k2p_grid.txt
