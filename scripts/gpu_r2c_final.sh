#!/bin/bash
# Round-2 third-session single-GPU evidence run on the final tree: full GPU suite (the reference's app/ staged, so the
# handler tests run), smoke, default bench, reference arm, encoder / serve benches, ncu launch list of the default bench.
# Outputs -> gpurun_out/r2f/
set -u
O=gpurun_out/r2f
mkdir -p $O
timeout 1200 python -m pytest tests -m gpu -q --timeout 600 2>&1 | tail -8 > $O/gpu_tests.log; tail -2 $O/gpu_tests.log
timeout 600 python -m pytest tests/test_reference_handlers_gpu.py tests/test_gpu_encoder.py -m gpu -q -s -k "reference or ollama" --timeout 500 2>&1 | grep -v "^PMC\|^$" | tail -30 > $O/reference_handlers.log; tail -2 $O/reference_handlers.log
timeout 300 python __graft_entry__.py smoke 2>&1 | tail -2 > $O/smoke.log; tail -1 $O/smoke.log
timeout 900 python bench.py > $O/bench_default.json 2> $O/bench_default.err; tail -c 300 $O/bench_default.json; echo
timeout 600 python bench.py --impl reference > $O/bench_reference.json 2> $O/bench_reference.err; tail -c 200 $O/bench_reference.json; echo
timeout 600 python bench.py --workload encode > $O/bench_encode.json 2> $O/bench_encode.err; tail -c 200 $O/bench_encode.json; echo
timeout 900 python bench.py --workload serve > $O/bench_serve.json 2> $O/bench_serve.err; tail -c 200 $O/bench_serve.json; echo
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 6000 --csv --log-file $O/launches_default.csv python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-traffic-probe > $O/ncu_default.log 2>&1; tail -1 $O/ncu_default.log | cut -c1-200
ls -la $O
