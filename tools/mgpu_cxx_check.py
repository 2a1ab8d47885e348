#!/usr/bin/env python
"""C++ multi-GPU driver at benchmark size (run on a box with N >= 2 GPUs):
  1. the twopaco CLI on the C2 FASTA files with TPC_GPUS=N vs TPC_GPUS=1: byte-identical files;
  2. tpc_multi_junctions_host (one process, N host threads, NCCL) on a workload (default C3) from pinned host buffers:
     image digest vs tests/golden/workload_digests.json, wall time per step.
Prints one JSON line.   python tools/mgpu_cxx_check.py [workload] [n_gpus]"""
import json
import os
import subprocess
import sys
import tempfile
import time
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT))
import bench  # noqa: E402
from oracle import oracle as O  # noqa: E402
from tools import benchutil  # noqa: E402
from twopaco_b200 import api  # noqa: E402


def main():
    import torch
    name = sys.argv[1] if len(sys.argv) > 1 else "c3"
    n = int(sys.argv[2]) if len(sys.argv) > 2 else torch.cuda.device_count()
    out = {"n_gpus": n}
    # ---- 1. CLI, C2 files
    wl = bench.WORKLOADS["c2"]
    dg = benchutil.synth_family_device(wl["seed"], wl["genomes"], wl["records"], wl["length"], wl["p"])
    with tempfile.TemporaryDirectory(prefix="tpc_mcli_") as d:
        paths = []
        for g in range(wl["genomes"]):
            p = os.path.join(d, f"g{g}.fa")
            O.write_fasta(p, [dg.record_ascii(g)], names=[f"g{g}_c0"])
            paths.append(p)
        del dg
        cli = str(ROOT / "twopaco_b200" / "bin" / "twopaco")
        digests, secs = [], []
        for gpus in (1, n):
            o = os.path.join(d, f"o{gpus}.bin")
            t0 = time.perf_counter()
            p = subprocess.run([cli, "-k", str(wl["k"]), "-f", str(wl["f"]), "-t", str(os.cpu_count()), "--tmpdir", d, "-o", o, *paths],
                               capture_output=True, text=True, env={**os.environ, "TPC_GPUS": str(gpus), "TPC_VERBOSE": "1"})
            secs.append(round(time.perf_counter() - t0, 3))
            assert p.returncode == 0, p.stderr
            digests.append(api.image_digest_host(open(o, "rb").read()))
            out[f"cli_c2_{gpus}gpu_breakdown"] = [ln for ln in p.stderr.splitlines() if ln.startswith("[tpc_build]")]
        gold = bench.golden_digest("c2")
        out["cli_c2_seconds"] = secs
        out["cli_c2_multi_equals_single"] = digests[0] == digests[1]
        out["cli_c2_equals_golden"] = [f"{digests[1][0]:016x}", f"{digests[1][1]:016x}"] == gold["digest"]
    # ---- 2. host buffers -> host buffers through tpc_multi_junctions_host
    wl = bench.WORKLOADS[name]
    dg = benchutil.synth_family_device(wl["seed"], wl["genomes"], wl["records"], wl["length"], wl["p"], keep_ascii=False)
    host = dg.to_host()
    total_bp = dg.total_bp
    dg.codes.close(); dg.n_mask.close()
    codes = torch.from_numpy(host.codes.view(np.int64)).pin_memory()
    nmask = torch.from_numpy(host.n_mask.view(np.int64)).pin_memory()
    pinned = api.PackedGenome(codes.numpy().view(np.uint64), nmask.numpy().view(np.uint64), host.n_positions, host.rec_start, host.rec_len)
    gold = bench.golden_digest(name)
    cap = (gold["records"] + len(host.rec_len)) * 12 + 4096 if gold else 8 << 30
    buf = torch.empty(cap, dtype=torch.uint8).pin_memory().numpy()
    mg = api.MultiGpu(n)
    times = []
    for i in range(4):
        t0 = time.perf_counter()
        img, st = mg.junctions_host(pinned, k=wl["k"], filter_bits=wl["f"], q=wl["q"], out=buf)
        if i:
            times.append(time.perf_counter() - t0)
    d = api.image_digest_host(img)
    mg.close()
    out.update({"workload": wl["name"], "host_to_host_ms": round(float(np.mean(times)) * 1e3, 2),
                "host_to_host_Gbps": round(total_bp / float(np.mean(times)) / 1e9, 3), "image_bytes": int(len(img)),
                "image_digest": [f"{d[0]:016x}", f"{d[1]:016x}"], "equals_golden": (None if not gold else gold["digest"] == [f"{d[0]:016x}", f"{d[1]:016x}"]),
                "stages_ms_max_over_shards": {k: round(getattr(st, k), 2) for k in ("ms_bin", "ms_fill", "ms_query", "ms_insert", "ms_index", "ms_emit", "ms_wall_candidates")},
                "junctions": st.junctions, "records": st.occurrences})
    print(json.dumps(out))


if __name__ == "__main__":
    main()
