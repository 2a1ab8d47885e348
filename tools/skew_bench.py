#!/usr/bin/env python
"""Repeat-rich input through the binned path: a 7-genome family (0.1 % divergence, k = 25, -f 32)
in which `frac` of every record is poly-A / (AC)n microsatellite runs, so that a handful of k-mers carry a few per cent of
all records and their filter slices overflow the uniform record arrays.  Timed: the automatic path (binned; re-binned
with exact per-slice capacities when the overflow list runs over), the direct kernels, and the same family without
repeats.  Images are compared (binned == direct).  Prints one JSON line.
    python tools/skew_bench.py [record_len] [frac]"""
import json
import os
import sys
import time
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT))
from tools.benchutil import hostsynth  # noqa: E402
from twopaco_b200 import api  # noqa: E402


def family(record_len, frac, seed=0x5EED):
    rng = np.random.default_rng(seed)
    recs = []
    for g in range(7):
        r = bytearray(hostsynth.record_prefix(seed, 1, 0.001, g, 0, record_len, record_len))
        if frac > 0:   # runs of 2000 bp every 2000 / frac bp, alternately poly-A and (AC)n, at the same founder sites in every genome
            step = int(2000 / frac)
            for i, at in enumerate(range(step // 2, len(r) - 2000, step)):
                r[at:at + 2000] = b"A" * 2000 if i % 2 == 0 else b"AC" * 1000
        recs.append(bytes(r))
    return recs


def timed(genome, **env):
    for k, v in env.items():
        os.environ[k] = v
    try:
        best, st, img = None, None, None
        for _ in range(3):
            t0 = time.perf_counter()
            img, st = api.junctions_host(genome, k=25, filter_bits=32, q=5)
            dt = time.perf_counter() - t0
            best = dt if best is None else min(best, dt)
        return best, st, api.image_digest_host(img)
    finally:
        for k in env:
            os.environ.pop(k, None)


def main():
    record_len = int(sys.argv[1]) if len(sys.argv) > 1 else 40_000_000
    frac = float(sys.argv[2]) if len(sys.argv) > 2 else 0.05
    out = {"record_len": record_len, "genomes": 7, "repeat_fraction": frac}
    for name, f in (("plain", 0.0), ("repeat_rich", frac)):
        g = api.pack_records(family(record_len, f))
        t_auto, st, d_auto = timed(g)
        t_dir, st_d, d_dir = timed(g, TPC_FILTER_MODE="direct")
        out[name] = {"bp": g.total_bp, "auto_ms": round(t_auto * 1e3, 1), "direct_ms": round(t_dir * 1e3, 1),
                     "auto_path": "binned" if st.bin_waves else "direct", "skew_rebins": st.skew_rebins,
                     "stages_ms": {k: round(getattr(st, k), 2) for k in ("ms_bin", "ms_fill", "ms_query", "ms_insert", "ms_emit")},
                     "junctions": st.junctions, "records": st.occurrences, "binned_equals_direct": d_auto == d_dir}
    print(json.dumps(out))


if __name__ == "__main__":
    main()
