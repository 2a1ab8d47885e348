#!/usr/bin/env python
"""Random-access roofline probe: sector touches/s into a 2^f-bit table."""
import json, sys
sys.path.insert(0, ".")
from twopaco_b200 import api
out = {}
for f in (int(a) for a in (sys.argv[1:] or ["32", "36"])):
    for mode, name in ((0, "load32B"), (1, "atomicOr"), (2, "load+condAtomicOr")):
        v = api.random_access_probe(f, mode, 1 << 31)
        out[f"f{f}_{name}"] = {"Gtouch/s": round(v / 1e9, 2), "GB/s@32B": round(v * 32 / 1e9, 1)}
print(json.dumps(out, indent=1))
