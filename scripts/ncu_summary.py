#!/usr/bin/env python
"""Summarise ncu outputs brought back in gpurun_out/ into small tracked files under profiles/.

    python scripts/ncu_summary.py full  gpurun_out/x.ncu-rep  profiles/r1_x_full.json
    python scripts/ncu_summary.py list  gpurun_out/launches.csv profiles/r1_x_launches.csv
"""
import csv
import io
import json
import subprocess
import sys

KEEP = [
    "Kernel Name", "launch__grid_size", "launch__block_size", "launch__registers_per_thread",
    "launch__shared_mem_per_block_dynamic", "launch__shared_mem_per_block_static",
    "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
    "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
    "sm__throughput.avg.pct_of_peak_sustained_elapsed",
    "sm__warps_active.avg.pct_of_peak_sustained_active",
    "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_tensor.sum", "sm__cycles_active.avg", "smsp__cycles_active.avg",
    "lts__t_sector_hit_rate.pct", "lts__t_bytes.sum", "l1tex__t_bytes.sum",
    "sm__cycles_elapsed.max", "smsp__inst_executed.sum",
    "smsp__warp_issue_stalled_long_scoreboard_per_warp_active.pct",
    "smsp__warp_issue_stalled_barrier_per_warp_active.pct",
]


def full(rep, out):
    txt = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    txt = txt[txt.index('"ID"'):]
    rows = list(csv.reader(io.StringIO(txt)))
    hdr, units = rows[0], rows[1]
    res = []
    for r in rows[2:]:
        d = {}
        for name in KEEP:
            if name in hdr:
                i = hdr.index(name)
                d[name] = r[i] + (" " + units[i] if units[i] else "")
        res.append(d)
    with open(out, "w") as f:
        json.dump({"source": rep, "tool": "ncu --set full --clock-control none", "launches": res}, f, indent=1)
    print(json.dumps(res, indent=1))


def launch_list(src, out):
    with open(src) as f:
        txt = f.read()
    txt = txt[txt.index('"ID"'):]
    rows = list(csv.DictReader(io.StringIO(txt)))
    total = sum(float(r["Metric Value"]) for r in rows)
    with open(out, "w") as f:
        f.write("id,kernel,grid,block,duration_ns,share_of_listed\n")
        for r in rows:
            name = r["Kernel Name"]
            short = name.split("(")[0][-90:]
            f.write(f'{r["ID"]},"{short}","{r["Grid Size"]}","{r["Block Size"]}",{r["Metric Value"]},'
                    f'{float(r["Metric Value"]) / total:.4f}\n')
    print("wrote", out, len(rows), "launches")


if __name__ == "__main__":
    {"full": full, "list": launch_list}[sys.argv[1]](sys.argv[2], sys.argv[3])
