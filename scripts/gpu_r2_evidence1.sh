#!/bin/bash
# Round 2 evidence on ONE GPU: full GPU test suite, driver-like bench lines, ncu launch list and full captures.
set -x
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -x -q -m gpu 2>&1 | tail -8 > gpurun_out/r2_final_gpu_tests.log; cat gpurun_out/r2_final_gpu_tests.log
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r2_final_smoke.log 2>&1; tail -1 gpurun_out/r2_final_smoke.log
python bench.py > gpurun_out/r2_final_bench_default.json 2> gpurun_out/r2_final_bench_default.err; tail -c 400 gpurun_out/r2_final_bench_default.json
python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/r2_final_bench_reference.json 2>&1; tail -c 300 gpurun_out/r2_final_bench_reference.json
B="python bench.py --no-cpu-baseline --no-traffic-probe"
timeout 600 $B --workload b1 --rows 1000000 --dtype fp32 > gpurun_out/r2_final_bench_b1_fp32_1m.json 2>&1
timeout 600 $B --workload b1 --rows 1000000 --dtype fp32 --prefilter > gpurun_out/r2_final_bench_b1_fp32_1m_prefilter.json 2>&1
timeout 600 $B --workload serve > gpurun_out/r2_final_bench_serve.json 2>&1; tail -c 300 gpurun_out/r2_final_bench_serve.json
timeout 600 $B --workload serve --no-prefilter > gpurun_out/r2_final_bench_serve_k2.json 2>&1
timeout 600 $B --workload b1024 --rows 12500000 --batch 256 --k 100 --dtype fp16 --no-cfg4 > gpurun_out/r2_final_bench_cfg4_shard.json 2>&1
timeout 600 $B --workload config1 > gpurun_out/r2_final_bench_config1.json 2>&1
# launch list of the default command (short)
python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-traffic-probe --no-cfg4 --no-yardstick > gpurun_out/plain.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r2_final_launches_default.csv \
    python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-traffic-probe --no-cfg4 --no-yardstick > gpurun_out/ncu_list.log 2>&1
# full captures: the int8 scan at the headline shape, K2 k=100 log mode at the configs[3] shard, K2 k=1, the b=1 scan
python scripts/k2p_probe.py 10000000 1024 10 bf16 > gpurun_out/plain2.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:topk_batched_i8 -s 4 -c 1 -f -o gpurun_out/r2_final_i8_10m python scripts/k2p_probe.py 10000000 1024 10 bf16 > gpurun_out/ncu_i8.log 2>&1
ncu --set full --clock-control none -k regex:batched_rescore -s 4 -c 1 -f -o gpurun_out/r2_final_rescore_10m python scripts/k2p_probe.py 10000000 1024 10 bf16 > gpurun_out/ncu_resc.log 2>&1
ncu --set full --clock-control none -k regex:topk_batched_kernel -s 2 -c 1 -f -o gpurun_out/r2_final_k2_cfg4 python bench.py --workload b1024 --rows 12500000 --batch 256 --k 100 --dtype fp16 --no-cfg4 --no-cpu-baseline --no-traffic-probe --no-e2e --no-sustained --no-yardstick --no-secondary --steps 3 --warmup 1 > gpurun_out/ncu_k2.log 2>&1
ncu --set full --clock-control none -k regex:topk_batched_i8 -s 4 -c 1 -f -o gpurun_out/r2_final_i8_cache64 python scripts/k2p_probe.py 1000000 64 1 bf16 > gpurun_out/ncu_c64.log 2>&1
ls -la gpurun_out/*.ncu-rep
