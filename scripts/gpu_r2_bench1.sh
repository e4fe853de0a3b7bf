#!/bin/bash
# Round 2: the default bench line on one GPU (what the driver runs), the reference arm, smoke.
set -x
mkdir -p gpurun_out
python bench.py > gpurun_out/r2_bench_default.json 2> gpurun_out/r2_bench_default.err
tail -c 6000 gpurun_out/r2_bench_default.json; tail -5 gpurun_out/r2_bench_default.err
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
