#!/bin/bash
# Round 2: programmatic dependent launch of the one-query scans.  1 GPU: parity + b1 numbers.
set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "fused or gemv or prefilter or graph or exchange or index or cache or stream" 2>&1 | tail -15 > gpurun_out/r2_pdl_tests.log
cat gpurun_out/r2_pdl_tests.log
timeout 600 python -m pytest tests/test_gpu_parity_at_size.py -x -q -m gpu -s -k "config2" 2>&1 | tail -5
B="python bench.py --no-cpu-baseline --no-cfg4 --no-traffic-probe --no-yardstick --no-e2e"
timeout 600 $B --steps 10 > gpurun_out/r2_pdl_default.json 2> gpurun_out/r2_pdl_default.err
python - <<'PY'
import json
l=json.loads(open("gpurun_out/r2_pdl_default.json").read().strip().splitlines()[-1])
print("b1024", round(l["value"]), l["roofline"]["kernel_ms"], l["roofline"].get("in_kernel"))
print("secondary", json.dumps(l["secondary"])[:900])
print("secondary_pf", json.dumps(l["secondary_prefiltered"])[:900])
PY
