#!/usr/bin/env python
"""Condenses an `ncu --set full` report into what profiles/ keeps: a per-kernel metric CSV and profiles/ncu_summary.json
(per-launch DRAM traffic etc., read by bench.py's roofline block).  Usage: tools/ncu_summary.py report.ncu-rep TAG"""
import csv
import io
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
KEEP = ("gpu__time_duration", "dram__bytes", "dram__throughput", "lts__t_sector_hit", "lts__t_bytes", "lts__throughput", "l1tex__t_sector_hit",
        "l1tex__throughput", "l1tex__t_bytes", "launch__", "sm__throughput", "sm__warps_active", "sm__cycles_active", "sm__cycles_elapsed",
        "sm__inst_executed.sum", "smsp__issue_active", "smsp__inst_executed.sum", "smsp__thread_inst_executed_per_inst", "smsp__average_warps_issue_stalled",
        "smsp__warps_eligible", "sass__inst_executed_local", "gpu__compute_memory_throughput", "sm__pipe_", "smsp__cycles_active")


def main(rep, tag):
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    hdr, units, data = rows[0], rows[1], rows[2:]
    name_i = hdr.index("Kernel Name")
    names = [r[name_i].split("(")[0].replace("void ", "") for r in data]
    keep = [i for i, h in enumerate(hdr) if h.startswith(KEEP)]
    out_csv = os.path.join(ROOT, "profiles", "%s_ncu_full_kernels.csv" % tag)
    with open(out_csv, "w", newline="") as f:
        w = csv.writer(f)
        w.writerow(["metric", "unit"] + ["%s#%s" % (n, r[0]) for n, r in zip(names, data)])
        for i in keep:
            w.writerow([hdr[i], units[i]] + [r[i] for r in data])

    def val(r, key, scale=1.0):
        i = hdr.index(key)
        u = units[i]
        mult = {"Mbyte": 1e6, "Kbyte": 1e3, "Gbyte": 1e9, "byte": 1.0}.get(u, 1.0)
        return float(r[i].replace(",", "")) * mult * scale

    summary = {"source": os.path.basename(rep), "tag": tag, "kernels": {}}
    for n, r in zip(names, data):
        base = n.split("<")[0]
        if base in summary["kernels"]:
            continue
        summary["kernels"][base] = {
            "name": n, "duration_us_under_ncu": val(r, "gpu__time_duration.sum") / (1e3 if units[hdr.index("gpu__time_duration.sum")] == "ns" else 1.0),
            "dram_bytes_per_launch": val(r, "dram__bytes_read.sum") + val(r, "dram__bytes_write.sum"),
            "dram_read_bytes": val(r, "dram__bytes_read.sum"), "dram_write_bytes": val(r, "dram__bytes_write.sum"),
            "registers_per_thread": val(r, "launch__registers_per_thread"), "grid": val(r, "launch__grid_size"),
            "issue_active_pct": val(r, "smsp__issue_active.avg.pct_of_peak_sustained_active"),
            "warps_active_pct": val(r, "sm__warps_active.avg.pct_of_peak_sustained_active"),
            "sm_cycles_active_avg": val(r, "sm__cycles_active.avg"), "sm_cycles_active_max": val(r, "sm__cycles_active.max"),
            "sm_cycles_elapsed_avg": val(r, "sm__cycles_elapsed.avg"),
            "warp_inst": val(r, "smsp__inst_executed.sum"), "l1_hit_pct": val(r, "l1tex__t_sector_hit_rate.pct"), "l2_hit_pct": val(r, "lts__t_sector_hit_rate.pct"),
            "stall_per_issue": {k.split("issue_stalled_")[1].replace("_per_issue_active.ratio", ""): round(val(r, k), 3)
                                for k in hdr if k.startswith("smsp__average_warps_issue_stalled_") and k.endswith("_per_issue_active.ratio")},
        }
    with open(os.path.join(ROOT, "profiles", "ncu_summary.json"), "w") as f:
        json.dump(summary, f, indent=1, sort_keys=True)
    for k, v in summary["kernels"].items():
        top = sorted(v["stall_per_issue"].items(), key=lambda kv: -kv[1])[:4]
        print("%-16s %.1f us  dram %.1f MB  issue %.0f%%  warps %.0f%%  regs %d  sm active avg/max/elapsed %.0fk/%.0fk/%.0fk  stalls %s" % (
            k, v["duration_us_under_ncu"], v["dram_bytes_per_launch"] / 1e6, v["issue_active_pct"], v["warps_active_pct"], v["registers_per_thread"],
            v["sm_cycles_active_avg"] / 1e3, v["sm_cycles_active_max"] / 1e3, v["sm_cycles_elapsed_avg"] / 1e3, top))


if __name__ == "__main__":
    main(sys.argv[1], sys.argv[2])
