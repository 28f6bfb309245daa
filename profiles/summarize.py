#!/usr/bin/env python
"""Turns the ncu artefacts brought back in gpurun_out/ into the small tracked summaries under profiles/.

  python profiles/summarize.py <tag> <launches.csv> <prof.ncu-rep> [kernel-regex]

writes profiles/<tag>_launches.csv (per-launch gpu__time_duration), profiles/<tag>_<kernel>_metrics.txt (the raw
metrics the roofline numbers come from), profiles/<tag>_<kernel>_stalls.txt (per-instruction stall samples, top 40)
and updates profiles/traffic_k_line.json (dram bytes per launch, read by bench.py for roofline.traffic).
"""
import csv
import io
import json
import os
import re
import subprocess
import sys

HERE = os.environ.get("SUMMARIZE_OUT") or os.path.dirname(os.path.abspath(__file__))  # (on the GPU box: gpurun_out/)
METRICS = [
    "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
    "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "dram__cycles_active.avg",
    "sm__warps_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread",
    "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem", "launch__grid_size",
    "launch__block_size", "launch__shared_mem_per_block_dynamic", "lts__t_sector_hit_rate.pct",
    "l1tex__t_sector_hit_rate.pct", "smsp__inst_executed.sum", "smsp__issue_active.avg.pct_of_peak_sustained_active",
    "sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
    "lts__t_bytes.sum", "smsp__inst_executed_op_local_ld.sum", "smsp__inst_executed_op_local_st.sum",
]


def ncu_csv(rep, page):
    out = subprocess.run(["ncu", "-i", rep, "--page", page, "--csv"], capture_output=True, text=True).stdout
    return list(csv.reader(io.StringIO(out)))


def main():
    tag, launches, rep = sys.argv[1:4]
    kernel = sys.argv[4] if len(sys.argv) > 4 else "k_line"
    # 1. launch list ("-": none for this capture)
    rows = [r for r in csv.reader(open(launches)) if len(r) > 10] if launches != "-" else []
    with open(os.path.join(HERE, f"{tag}_launches.csv") if rows else os.devnull, "w") as f:
        f.write("id,kernel,grid,block,gpu__time_duration_ns\n")
        total = {}
        for r in rows[1:]:
            name = re.sub(r"smore::|<unnamed>::", "", r[4])
            f.write(f"{r[0]},\"{name}\",\"{r[8]}\",\"{r[7]}\",{r[-1]}\n")
            short = name.split("<")[0].replace("void ", "")
            total[short] = total.get(short, 0) + float(r[-1])
        s = sum(total.values())
        f.write("# share of device time per kernel (cold-cache, serialised under ncu: compare shares, not absolutes)\n")
        for k, v in sorted(total.items(), key=lambda x: -x[1]):
            f.write(f"# {k}: {v / 1e6:.3f} ms ({100 * v / s:.1f}%)\n")
    # 2. raw metrics of the profiled kernel
    raw = ncu_csv(rep, "raw")
    hdr, units, data = raw[0], raw[1], raw[2:]
    with open(os.path.join(HERE, f"{tag}_{kernel}_metrics.txt"), "w") as f:
        f.write(f"# ncu --set full --clock-control none, kernel {data[0][hdr.index('Kernel Name')]}\n")
        for m in METRICS:
            if m in hdr:
                i = hdr.index(m)
                f.write(f"{m} [{units[i]}] = {', '.join(d[i] for d in data)}\n")
    if "dram__bytes_read.sum" in hdr:
        def gb(m):
            i = hdr.index(m)
            scale = {"Gbyte": 1e9, "Mbyte": 1e6, "Kbyte": 1e3, "byte": 1.0}[units[i]]
            return float(data[0][i]) * scale
        traffic = gb("dram__bytes_read.sum") + gb("dram__bytes_write.sum")
        if kernel == "k_line":
            json.dump({"dram_bytes_per_launch": traffic, "source": f"profiles/{tag}_{kernel}_metrics.txt",
                       "note": "dram__bytes_read.sum + dram__bytes_write.sum of one k_line launch of 2^24 updates"},
                      open(os.path.join(HERE, "traffic_k_line.json"), "w"))
    # 3. stall samples per SASS instruction
    src = ncu_csv(rep, "source")
    his = [i for i, r in enumerate(src) if r and r[0] == "Address"]
    hi = his[0]
    end = his[1] - 1 if len(his) > 1 else len(src)
    h = src[hi]
    idx = {n: i for i, n in enumerate(h)}
    body = [r for r in src[hi + 1:end] if len(r) == len(h)]

    def I(x):
        try:
            return int(x)
        except ValueError:
            return 0
    tot = sum(I(r[idx["# Samples"]]) for r in body)
    stalls = [n for n in h if n.startswith("stall_") and "Not Issued" not in n]
    with open(os.path.join(HERE, f"{tag}_{kernel}_stalls.txt"), "w") as f:
        f.write(f"# warp-stall samples, total {tot}\n")
        for k, v in sorted(((s, sum(I(r[idx[s]]) for r in body)) for s in stalls), key=lambda x: -x[1])[:8]:
            f.write(f"{k}: {v} ({100 * v / max(tot, 1):.1f}%)\n")
        f.write("# top instructions: samples | dominant stall | SASS\n")
        for r in sorted(body, key=lambda r: -I(r[idx["# Samples"]]))[:40]:
            dom = max(stalls, key=lambda s: I(r[idx[s]]))
            f.write(f"{r[idx['# Samples']]} | {dom} | {r[idx['Source']]}\n")


if __name__ == "__main__":
    main()
