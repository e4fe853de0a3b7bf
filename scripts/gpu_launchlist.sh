set -x
cd /root/repo
O=gpurun_out/evidence2; mkdir -p $O
CMD="python bench.py --no-cpu-baseline"
timeout 600 $CMD > $O/plain_default.json 2> $O/plain_default.err &&
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file $O/launches_default.csv $CMD > $O/ncu_list.log 2>&1
CMD2="python bench.py --workload b1 --rows 2000000 --steps 3 --warmup 1 --no-cpu-baseline --no-e2e"
timeout 300 $CMD2 > $O/plain_b1.log 2>&1 &&
timeout 600 ncu --set full --clock-control none --import-source on -k regex:topk_gemv_kernel -s 2 -c 1 -o $O/gemv $CMD2 > $O/ncu_gemv.log 2>&1
tail -2 $O/ncu_gemv.log
