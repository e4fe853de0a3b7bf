#!/bin/bash
# Last verification of the round on the final tree: full GPU suite, smoke, default bench line.  Outputs -> gpurun_out/r2g/
set -u
O=gpurun_out/r2g
mkdir -p $O
timeout 1200 python -m pytest tests -m gpu -q --timeout 600 2>&1 | tail -8 > $O/gpu_tests.log; tail -2 $O/gpu_tests.log
timeout 300 python __graft_entry__.py smoke 2>&1 | tail -2 > $O/smoke.log; tail -1 $O/smoke.log
timeout 900 python bench.py > $O/bench_default.json 2> $O/bench_default.err; tail -c 300 $O/bench_default.json; echo
