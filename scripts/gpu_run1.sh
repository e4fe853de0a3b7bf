set -x
nvidia-smi --query-gpu=name,memory.total,clocks.max.sm --format=csv
python -m pytest tests -m gpu -x -q 2>&1 | tail -40 > gpurun_out/r1_tests.log
cat gpurun_out/r1_tests.log | tail -15
python bench.py --workload b1 --steps 30 --warmup 3 > gpurun_out/r1_bench_b1_bf16.json 2> gpurun_out/r1_bench_b1_bf16.err; tail -3 gpurun_out/r1_bench_b1_bf16.err; cat gpurun_out/r1_bench_b1_bf16.json
python bench.py --workload b1 --rows 1000000 --dtype fp32 --steps 50 --warmup 5 --no-cpu-baseline > gpurun_out/r1_bench_b1_fp32_1m.json 2> gpurun_out/r1_bench_b1_fp32.err; cat gpurun_out/r1_bench_b1_fp32_1m.json
python bench.py --workload b1 --rows 2000000 --steps 3 --warmup 1 --no-cpu-baseline --no-e2e > gpurun_out/plain.log 2>&1 && ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r1_launches_b1.csv python bench.py --workload b1 --rows 2000000 --steps 3 --warmup 1 --no-cpu-baseline --no-e2e > gpurun_out/ncu1.log 2>&1
python bench.py --workload b1 --rows 2000000 --steps 3 --warmup 1 --no-cpu-baseline --no-e2e > gpurun_out/plain.log 2>&1 && ncu --set full --clock-control none --import-source on -k regex:topk_gemv -s 2 -c 2 -o gpurun_out/r1_gemv python bench.py --workload b1 --rows 2000000 --steps 3 --warmup 1 --no-cpu-baseline --no-e2e > gpurun_out/ncu2.log 2>&1
tail -5 gpurun_out/ncu2.log
