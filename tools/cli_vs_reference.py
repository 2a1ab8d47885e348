#!/usr/bin/env python
"""Files -> file, full size: our `twopaco` CLI vs the UNMODIFIED reference `twopaco -t $(nproc)` on the
same FASTA files of a BASELINE config (default C2: 62 x 5 Mbp, k=25, -f 32, -q 5), outputs compared
through the canonical relabelling.  Prints one JSON line.  Run on the GPU box:
    python tools/cli_vs_reference.py [c2|dev]
"""
import json
import os
import subprocess
import sys
import tempfile
import time
from pathlib import Path

ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT))
from bench import WORKLOADS  # noqa: E402
from oracle import oracle as O  # noqa: E402
from tools import benchutil  # noqa: E402
from twopaco_b200 import api  # noqa: E402


def main():
    wl = WORKLOADS[sys.argv[1] if len(sys.argv) > 1 else "c2"]
    cores = os.cpu_count() or 1
    dg = benchutil.synth_family_device(wl["seed"], wl["genomes"], wl["records"], wl["length"], wl["p"])
    with tempfile.TemporaryDirectory(prefix="tpc_cli_") as d:
        paths = []
        for g in range(wl["genomes"]):
            recs = [dg.record_ascii(g * wl["records"] + c) for c in range(wl["records"])]
            p = os.path.join(d, f"g{g}.fa")
            O.write_fasta(p, recs, names=[f"g{g}_c{c}" for c in range(wl["records"])])
            paths.append(p)
        total_bp = dg.total_bp
        del dg
        cli = str(ROOT / "twopaco_b200" / "bin" / "twopaco")
        ours = os.path.join(d, "ours.bin")
        t0 = time.perf_counter()
        p = subprocess.run([cli, "-k", str(wl["k"]), "-f", str(wl["f"]), "-q", str(wl["q"]), "-t", str(cores), "--tmpdir", d,
                            "-o", ours, *paths], capture_output=True, text=True)
        t_ours = time.perf_counter() - t0
        assert p.returncode == 0, p.stderr
        t0 = time.perf_counter()
        p2 = subprocess.run([cli, "-k", str(wl["k"]), "-f", str(wl["f"]), "-q", str(wl["q"]), "-t", str(cores), "--tmpdir", d,
                             "-o", ours + ".2", *paths], capture_output=True, text=True, env={**os.environ, "TPC_VERBOSE": "1"})
        t_ours_warm = min(t_ours, time.perf_counter() - t0)
        breakdown = [ln for ln in p2.stderr.splitlines() if ln.startswith("[tpc_build]")]
        os.remove(ours + ".2")
        ref = os.path.join(d, "ref.bin")
        t0 = time.perf_counter()
        r = subprocess.run([str(O.REF_TWOPACO), "-k", str(wl["k"]), "-f", str(wl["f"]), "-q", str(wl["q"]), "-t", str(cores),
                            "--tmpdir", d, "-o", ref, *paths], capture_output=True, text=True)
        t_ref = time.perf_counter() - t0
        assert r.returncode == 0, r.stderr
        a, b = open(ours, "rb").read(), open(ref, "rb").read()
        # canonical relabelling on the GPU (tpc_canonical_image_device): byte-identical canonical images <=> same graph
        ca, na = api.canonical_image(a)
        cb, nb = api.canonical_image(b)
        same = ca == cb and na == nb
        d = api.image_digest_host(a)
        try:
            import bench
            gold = bench.golden_digest(sys.argv[1] if len(sys.argv) > 1 else "c2")
        except Exception:
            gold = None
        dj = lambda s: [ln for ln in s.splitlines() if ln.startswith("Distinct junctions")]
        print(json.dumps({"workload": wl["name"], "total_bp": total_bp, "host_cores": cores,
                          "ours_cli_s": round(t_ours, 3), "ours_cli_best_of_2_s": round(t_ours_warm, 3),
                          "reference_cli_s": round(t_ref, 3), "speedup_files_to_file": round(t_ref / t_ours_warm, 1),
                          "ours_Gbps": round(total_bp / t_ours_warm / 1e9, 3), "reference_Gbps": round(total_bp / t_ref / 1e9, 5),
                          "image_bytes": [len(a), len(b)], "canonical_streams_identical": bool(same), "distinct_ids": [na, nb],
                          "ours_image_digest": [f"{d[0]:016x}", f"{d[1]:016x}"],
                          "ours_image_digest_equals_bench_golden": (None if not gold else gold["digest"] == [f"{d[0]:016x}", f"{d[1]:016x}"]),
                          "ours_breakdown": breakdown, "ours_log": dj(p.stdout), "reference_log": dj(r.stdout)}))


if __name__ == "__main__":
    main()
