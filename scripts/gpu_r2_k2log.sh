#!/bin/bash
# Round 2: K2 log mode (k > 32).  Parity first, then the BASELINE configs[3] shard shape.
set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "batched or adversarial or cache or random_call or full_size" 2>&1 | tail -15 > gpurun_out/r2_k2log_tests.log
cat gpurun_out/r2_k2log_tests.log
timeout 600 python -m pytest tests/test_gpu_parity_at_size.py -x -q -m gpu -s -k "config4 or config3" 2>&1 | tail -8 >> gpurun_out/r2_k2log_tests.log
tail -8 gpurun_out/r2_k2log_tests.log
B="python bench.py --no-cpu-baseline --no-e2e --no-secondary --no-cfg4 --no-sustained --no-traffic-probe --no-yardstick"
for K in 10 100 128; do
  timeout 300 $B --workload b1024 --rows 12500000 --batch 256 --k $K --dtype fp16 --steps 20 > gpurun_out/r2_k2log_shard_k$K.json 2> gpurun_out/r2_k2log_shard_k$K.err
  python - <<PY
import json
l=json.loads(open("gpurun_out/r2_k2log_shard_k$K.json").read().strip().splitlines()[-1])
r=l["roofline"]
print("k=$K value", round(l["value"]), "kernel_ms", round(r["kernel_ms"],3), "TF", round(r["achieved"]), "in_kernel", r.get("in_kernel"))
PY
done
timeout 300 $B --workload b1024 --rows 2500000 --k 100 --steps 20 > gpurun_out/r2_k2log_b1024_k100.json 2>&1
tail -c 600 gpurun_out/r2_k2log_b1024_k100.json
