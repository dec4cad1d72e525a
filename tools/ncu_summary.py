#!/usr/bin/env python
"""Summarise ncu output for profiles/: a launch list (csv from `--metrics gpu__time_duration.sum`)
and/or a full capture (.ncu-rep).  Usage:
    python tools/ncu_summary.py --launches gpurun_out/launches.csv --rep gpurun_out/prof.ncu-rep > profiles/NAME.txt
"""
import argparse
import csv
import subprocess
import sys
from collections import defaultdict

KEYS = [
    "gpu__time_duration.sum", "launch__grid_size", "launch__block_size", "launch__registers_per_thread",
    "launch__shared_mem_per_block_dynamic", "launch__occupancy_limit_shared_mem", "launch__occupancy_limit_registers",
    "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
    "sm__throughput.avg.pct_of_peak_sustained_elapsed", "l1tex__throughput.avg.pct_of_peak_sustained_elapsed",
    "lts__throughput.avg.pct_of_peak_sustained_elapsed",
    "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__issue_active.avg.pct_of_peak_sustained_active",
    "smsp__inst_executed.sum", "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active",
    "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active",
    "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
    "smsp__cycles_active.avg", "sm__cycles_elapsed.max",
    "l1tex__data_pipe_tc_wavefronts_mem_shared.sum", "smsp__inst_executed_pipe_fma.sum", "smsp__inst_executed_pipe_alu.sum",
    "smsp__inst_executed_pipe_lsu.sum", "smsp__inst_executed_pipe_xu.sum", "sm__inst_executed_pipe_uniform.sum",
]


def launches(path):
    rows = [r for r in csv.reader(open(path)) if len(r) > 5]
    hdr = rows[0]
    ki, vi, ui = hdr.index("Kernel Name"), hdr.index("Metric Value"), hdr.index("Metric Unit")
    d = defaultdict(list)
    for r in rows[1:]:
        try:
            v = float(r[vi].replace(",", ""))
        except ValueError:
            continue
        scale = {"ns": 1e-3, "us": 1.0, "ms": 1e3, "s": 1e6}.get(r[ui], 1e-3)
        d[r[ki]].append(v * scale)
    tot = sum(sum(v) for v in d.values())
    print(f"# launch list: {path}  (cold-cache, serialised: compare SHARES)")
    for k, v in sorted(d.items(), key=lambda kv: -sum(kv[1])):
        print(f"{sum(v):12.1f} us  share {sum(v) / tot:6.3f}  n={len(v):4d}  avg {sum(v) / len(v):10.1f} us  {k[:110]}")


def rep(path):
    out = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    if len(rows) < 3:
        print("# could not read", path)
        return
    h, units = rows[0], rows[1]
    for launch in rows[2:]:
        name = launch[h.index("Kernel Name")] if "Kernel Name" in h else "?"
        print(f"# full capture: {path}  kernel: {name[:120]}")
        for k in KEYS:
            if k in h:
                i = h.index(k)
                print(f"{k:75s} {launch[i]:>18s} {units[i]}")


def traffic(path, key, utterances, out_path, source):
    """dram__bytes_read.sum + dram__bytes_write.sum of the capture's launches (averaged), per utterance ->
    out_path[key] (the file bench.py reads `roofline.traffic` from), stamped with the current commit."""
    import json
    import os
    out = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    h, units = rows[0], rows[1]
    ir, iw = h.index("dram__bytes_read.sum"), h.index("dram__bytes_write.sum")

    def to_bytes(v, u):
        mult = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}.get(u, 1.0)
        return float(v.replace(",", "")) * mult
    tot = [to_bytes(r[ir], units[ir]) + to_bytes(r[iw], units[iw]) for r in rows[2:]]
    commit = subprocess.run(["git", "rev-parse", "--short", "HEAD"], capture_output=True, text=True).stdout.strip()
    rec = {}
    if os.path.exists(out_path):
        rec = json.load(open(out_path))
    rec[key] = {"dram_bytes_per_utt": sum(tot) / len(tot) / utterances, "utterances": utterances, "launches_averaged": len(tot),
                "commit": commit, "source": source or path}
    json.dump(rec, open(out_path, "w"), indent=1)
    print(f"# traffic: {key}: {rec[key]}")


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--launches")
    ap.add_argument("--rep")
    ap.add_argument("--traffic-key", help="kernel name as bench.py reports it (roofline.dominant_kernel.kernel)")
    ap.add_argument("--utterances", type=int, help="utterances per captured launch")
    ap.add_argument("--traffic-out", default="profiles/r2_traffic.json")
    ap.add_argument("--source", default=None)
    a = ap.parse_args()
    if a.launches:
        launches(a.launches)
    if a.rep:
        rep(a.rep)
    if a.rep and a.traffic_key:
        traffic(a.rep, a.traffic_key, a.utterances, a.traffic_out, a.source)


if __name__ == "__main__":
    sys.exit(main())
