#!/bin/bash
# first contact of the CTA-pair (cta_group::2) kernel
set -x
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 120 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "ragged_batches and cta_pair" 2>&1 | tail -30 > gpurun_out/k2p_first.log
cat gpurun_out/k2p_first.log
timeout 600 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "batched or tensor_path or full_size" 2>&1 | tail -30 > gpurun_out/k2p_tests.log
cat gpurun_out/k2p_tests.log
for cg in 1 2; do
timeout 300 python bench.py --workload b1024 --rows 2000000 --steps 10 --warmup 3 --no-cpu-baseline --no-e2e --k2-cta-group $cg > gpurun_out/k2p_bench_2m_cg$cg.json 2> gpurun_out/k2p_bench_2m_cg$cg.err
tail -3 gpurun_out/k2p_bench_2m_cg$cg.err; cat gpurun_out/k2p_bench_2m_cg$cg.json
done
