#!/usr/bin/env python
"""Roofline probes: random sector touches/s into a 2^f-bit table in HBM, and inside one
L2-resident filter slice with the record stream beside it (the pattern of the binned apply kernels).
    python tools/probe.py [hbm f ...] | [slice]"""
import json
import sys
from pathlib import Path

sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
from tools import benchutil  # noqa: E402

out = {}
args = sys.argv[1:] or ["hbm", "32", "36", "slice"]
if "hbm" in args:
    for f in [int(a) for a in args if a.isdigit()] or [36]:
        for mode, name in ((0, "load32B"), (1, "atomicOr"), (2, "load+condAtomicOr")):
            v = benchutil.random_access_probe(f, mode, 1 << 31)
            out[f"hbm_f{f}_{name}"] = {"Gtouch/s": round(v / 1e9, 2), "GB/s@32B": round(v * 32 / 1e9, 1)}
if "slice" in args:
    for lg in (24, 25, 26, 27):
        for mode, name in ((0, "load"), (1, "query"), (2, "fill")):
            for U in (4, 8):
                for ctas in (1, 2, 4, 8):
                    if U == 8 and ctas == 8:
                        continue
                    v = benchutil.slice_probe(lg, 6, 32 << 20, 7, mode, U, ctas)
                    out[f"slice{1 << (lg - 20)}MiB_{name}_U{U}_ctas{ctas}"] = round(v / 1e9, 2)
print(json.dumps(out, indent=1))
