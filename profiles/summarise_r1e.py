#!/usr/bin/env python
"""Turns the r1e ncu outputs in gpurun_out/ into the text summaries kept under profiles/ (run in the build container:
`ncu -i` reads reports without a GPU)."""
import csv
import io
import json
import os
import subprocess
import sys
from collections import defaultdict

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
OUT = os.path.join(ROOT, "gpurun_out")
PROF = os.path.join(ROOT, "profiles")

KEYS = [
    "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
    "dram__throughput.avg.pct_of_peak_sustained_elapsed", "l1tex__throughput.avg.pct_of_peak_sustained_elapsed",
    "lts__throughput.avg.pct_of_peak_sustained_elapsed", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
    "smsp__issue_active.avg.pct_of_peak_sustained_active", "sm__warps_active.avg.pct_of_peak_sustained_active",
    "sm__pipe_tensor_subpipe_hmma_cycles_active.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_tensor.sum", "launch__grid_size", "launch__block_size", "launch__registers_per_thread",
    "launch__shared_mem_per_block_dynamic", "launch__occupancy_limit_shared_mem",
]


def raw_rows(rep):
    txt = subprocess.check_output(["ncu", "-i", rep, "--page", "raw", "--csv"], text=True)
    rows = list(csv.reader(io.StringIO(txt)))
    hdr, units = rows[0], rows[1]
    return hdr, units, rows[2:]


def num(x):
    try:
        return float(x.replace(",", ""))
    except ValueError:
        return None


def summarise(rep, out_name, title):
    hdr, units, rows = raw_rows(rep)
    unit = dict(zip(hdr, units))
    lines = [f"# {title}", f"# source: ncu --set full --clock-control none (report {os.path.basename(rep)}, read with ncu -i --page raw --csv)", ""]
    traffic = {}
    for r in rows:
        d = dict(zip(hdr, r))
        name = d["Kernel Name"]
        lines.append(f"== {name}   grid {d.get('launch__grid_size')} x block {d.get('launch__block_size')}")
        for k in KEYS:
            if k in d and d[k] != "":
                lines.append(f"   {k:78s} {d[k]:>16s} {unit.get(k, '')}")
        tens = [k for k in hdr if "tensor" in k and "pct" in k and d.get(k)]
        for k in tens[:6]:
            if k not in KEYS:
                lines.append(f"   {k:78s} {d[k]:>16s} {unit.get(k, '')}")
        stalls = [(k, num(d[k])) for k in hdr if k.startswith("smsp__average_warps_issue_stalled") and k.endswith("_per_issue_active.ratio")]
        stalls = sorted([(k, v) for k, v in stalls if v is not None], key=lambda kv: -kv[1])[:6]
        lines.append("   top stall reasons (warps stalled per issue-active cycle):")
        for k, v in stalls:
            lines.append(f"      {k.replace('smsp__average_warps_issue_stalled_', '').replace('_per_issue_active.ratio', ''):28s} {v:8.2f}")
        rd, wr = num(d.get("dram__bytes_read.sum", "")), num(d.get("dram__bytes_write.sum", ""))
        if rd is not None and wr is not None:
            scale = {"Gbyte": 1e9, "Mbyte": 1e6, "Kbyte": 1e3, "byte": 1.0}
            tot = rd * scale.get(unit.get("dram__bytes_read.sum"), 1.0) + wr * scale.get(unit.get("dram__bytes_write.sum"), 1.0)
            lines.append(f"   dram bytes read + write per launch: {tot / 1e6:.1f} MB")
            traffic.setdefault(name.split("(")[0].split("::")[-1].split("<")[0], []).append(tot)
        lines.append("")
    with open(os.path.join(PROF, out_name), "w") as f:
        f.write("\n".join(lines))
    return traffic


def launches(csv_path, out_name):
    rows = [r for r in csv.reader(open(csv_path)) if len(r) > 10 and r[0].isdigit()]
    per = defaultdict(list)
    for r in rows:
        name = r[4].split("(")[0]
        per[name].append(float(r[-1].replace(",", "")))
    unit = "ns"
    lines = ["# launch list summary of `python bench.py --steps 2 --warmup 1 --no-learner` under",
             "# ncu --metrics gpu__time_duration.sum --clock-control none -c 400 (cold-cache, serialised: compare shares)", ""]
    tot = sum(sum(v) for v in per.values())
    for name, v in sorted(per.items(), key=lambda kv: -sum(kv[1])):
        lines.append(f"{sum(v) / tot * 100:6.2f} %  n={len(v):4d}  mean {sum(v) / len(v) / 1e3:10.2f} us  max {max(v) / 1e3:10.2f} us  {name}")
    with open(os.path.join(PROF, out_name), "w") as f:
        f.write("\n".join(lines) + "\n")


if __name__ == "__main__":
    t = summarise(os.path.join(OUT, "r1e_env_gae.ncu-rep"), "r1e_ncu_env_gae_summary.txt",
                  "r1e: checkers_rollout_kernel + gae_tile_kernel of the bench step (16 384 envs x T = 128)")
    summarise(os.path.join(OUT, "r1e_learner.ncu-rep"), "r1e_ncu_learner_kernels_summary.txt",
              "r1e: GRU recurrence tile kernels (101 steps x 8 192 rows x 64) and the tensor-core mixer (409 600 rows)")
    launches(os.path.join(OUT, "r1e_launches_bench_steps2.csv"), "r1e_launches_bench_steps2_summary.txt")
    print({k: [round(x) for x in v] for k, v in t.items()})
