#!/bin/bash
set -x
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 900 python -m pytest tests -x -q -m gpu 2>&1 | tail -4
timeout 600 python bench.py > gpurun_out/r1m_bench_default.json 2> gpurun_out/r1m_bench_default.err; tail -2 gpurun_out/r1m_bench_default.err; cat gpurun_out/r1m_bench_default.json
