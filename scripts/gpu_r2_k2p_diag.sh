#!/bin/bash
set -x
mkdir -p gpurun_out
( timeout 300 python scripts/k2p_modes.py 2500000 1024 10
  timeout 300 python scripts/k2p_modes.py 1000000 64 1
  timeout 300 python scripts/k2p_modes.py 10000000 128 10 ) 2>&1 | grep "rows=" | tee gpurun_out/r2_k2p_modes.txt
for args in "2500000 1024 10 bf16" "1000000 64 1 bf16"; do
  tag=$(echo $args | tr ' ' '_')
  timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv --log-file gpurun_out/r2_k2p_launches_$tag.csv python scripts/k2p_probe.py $args > /dev/null 2>&1
  python - <<PY
import csv
rows=[r for r in csv.reader(open("gpurun_out/r2_k2p_launches_$tag.csv")) if len(r)>10]
hdr=rows[0]; ik=hdr.index("Kernel Name"); iv=hdr.index("Metric Value")
agg={}
for r in rows[1:]:
    k=r[ik].split("(")[0][:60]; agg.setdefault(k,[]).append(float(r[iv].replace(",","")))
for k,v in agg.items(): print("$tag", k, len(v), "launches, median us", sorted(v)[len(v)//2]/1e3)
PY
done
