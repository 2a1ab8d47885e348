#!/usr/bin/env python
"""Summarise an `ncu --page raw --csv` export: one markdown row per kernel launch group and the
per-launch DRAM traffic (dram__bytes_read.sum + dram__bytes_write.sum) as JSON.

    ncu -i X.ncu-rep --page raw --csv > X_raw.csv ; python tools/ncu_summary.py X_raw.csv [out.md] [traffic.json]
"""
import collections
import csv
import json
import sys

COLS = ["launch__grid_size", "launch__registers_per_thread", "gpu__time_duration.sum", "dram__bytes_read.sum",
        "dram__bytes_write.sum", "lts__t_sectors.sum", "lts__t_sector_hit_rate.pct",
        "lts__throughput.avg.pct_of_peak_sustained_elapsed", "dram__throughput.avg.pct_of_peak_sustained_elapsed",
        "smsp__issue_active.avg.pct_of_peak_sustained_active", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "smsp__inst_executed.sum", "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio"]


def to_bytes(v, unit):
    f = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "Tbyte": 1e12}[unit]
    return float(v.replace(",", "")) * f


def to_ms(v, unit):
    f = {"ns": 1e-6, "us": 1e-3, "ms": 1, "s": 1e3}.get(unit, {"nsecond": 1e-6, "usecond": 1e-3, "msecond": 1, "second": 1e3}.get(unit, 1))
    return float(v.replace(",", "")) * f


def main():
    rows = list(csv.reader(open(sys.argv[1])))
    hdr, units = rows[0], rows[1]
    ki = hdr.index("Kernel Name")
    groups = collections.OrderedDict()
    for r in rows[2:]:
        name = r[ki].split("(")[0].replace("void ", "").replace("tpc::", "")
        groups.setdefault(name, []).append(r)
    cols = [c for c in COLS if c in hdr]
    lines = ["| kernel | launches | " + " | ".join(c.replace("__", " ").replace(".sum", "").replace(".avg.pct_of_peak_sustained_", " % ") for c in cols) + " |",
             "|---|---|" + "---|" * len(cols), "| | | " + " | ".join(units[hdr.index(c)] for c in cols) + " |"]
    traffic = {}
    for name, rs in groups.items():
        mid = rs[len(rs) // 2]
        lines.append(f"| {name} | {len(rs)} | " + " | ".join(mid[hdr.index(c)] for c in cols) + " |")
        ir, iw, it = hdr.index("dram__bytes_read.sum"), hdr.index("dram__bytes_write.sum"), hdr.index("gpu__time_duration.sum")
        b = [to_bytes(r[ir], units[ir]) + to_bytes(r[iw], units[iw]) for r in rs]
        t = [to_ms(r[it], units[it]) for r in rs]
        traffic[name] = {"launches_captured": len(rs), "avg_ms": round(sum(t) / len(t), 4), "dram_bytes_per_launch": round(sum(b) / len(b))}
    md = "\n".join(lines) + "\n"
    if len(sys.argv) > 2:
        open(sys.argv[2], "a").write(md)
    else:
        print(md)
    if len(sys.argv) > 3:
        json.dump(traffic, open(sys.argv[3], "w"), indent=1)
    else:
        print(json.dumps(traffic, indent=1))


if __name__ == "__main__":
    main()
