#!/bin/bash
# Round-2 (second session) single-GPU evidence run: full GPU suite, smoke, default bench, reference arm,
# encoder / serve benches, ncu --set full of one encoder block, ncu launch lists.  Outputs -> gpurun_out/r2b/
set -u
O=gpurun_out/r2b
mkdir -p $O
timeout 900 python -m pytest tests -m gpu -q --timeout 600 2>&1 | tail -8 > $O/gpu_tests.log; tail -2 $O/gpu_tests.log
timeout 300 python __graft_entry__.py smoke 2>&1 | tail -2 > $O/smoke.log; tail -1 $O/smoke.log
timeout 900 python bench.py > $O/bench_default.json 2> $O/bench_default.err; tail -c 300 $O/bench_default.json; echo
timeout 600 python bench.py --impl reference > $O/bench_reference.json 2> $O/bench_reference.err; tail -c 200 $O/bench_reference.json; echo
timeout 600 python bench.py --workload encode > $O/bench_encode.json 2> $O/bench_encode.err; tail -c 200 $O/bench_encode.json; echo
timeout 900 python bench.py --workload serve > $O/bench_serve.json 2> $O/bench_serve.err; tail -c 200 $O/bench_serve.json; echo
timeout 600 ncu --set full --clock-control none --import-source on -k regex:encoder_ -s 9 -c 6 -f -o $O/encoder_block python scripts/enc_profile.py > $O/ncu_block.log 2>&1; tail -2 $O/ncu_block.log
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 2000 --csv --log-file $O/launches_encode.csv python bench.py --workload encode --steps 2 --warmup 1 --no-cpu-baseline --no-e2e --no-yardstick > $O/ncu_encode.log 2>&1; tail -1 $O/ncu_encode.log | cut -c1-200
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 6000 --csv --log-file $O/launches_default.csv python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-traffic-probe > $O/ncu_default.log 2>&1; tail -1 $O/ncu_default.log | cut -c1-200
ls -la $O
