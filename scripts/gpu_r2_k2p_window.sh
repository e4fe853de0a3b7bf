#!/bin/bash
mkdir -p gpurun_out
( for w in -1 4 8 16 32; do
    K2P_WINDOW=$w timeout 300 python scripts/k2p_probe.py 10000000 1024 10 bf16 2>&1 | tail -1
  done
  K2P_WINDOW=-1 timeout 300 python scripts/k2p_probe.py 10000000 256 10 bf16 2>&1 | tail -1
  K2P_WINDOW=8 timeout 300 python scripts/k2p_probe.py 10000000 256 10 bf16 2>&1 | tail -1 ) | tee gpurun_out/r2_k2p_window.txt
for w in -1 8; do
  K2P_WINDOW=$w timeout 600 ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum --clock-control none -k regex:topk_batched_i8 -s 4 -c 1 --csv --log-file gpurun_out/r2_k2p_dram_w$w.csv python scripts/k2p_probe.py 10000000 1024 10 bf16 > /dev/null 2>&1
  grep -E "dram__bytes|gpu__time" gpurun_out/r2_k2p_dram_w$w.csv | awk -F'","' -v w=$w '{print "window=" w, $(NF-2), $(NF-1), $NF}'
done | tee -a gpurun_out/r2_k2p_window.txt
