#!/usr/bin/env python
"""Generate tests/golden/gfa_golden.json: md5 and size of what the UNMODIFIED reference graphdump (oracle/_ref/graphdump)
prints for -f gfa1 / gfa2 / fasta (with and without --prefix) on the image of every GFA case.

The image is the C oracle's (oracle/junction_oracle.c): its ids are the deterministic first-appearance numbering the CUDA
path also produces (the GPU parity tests require the two images to be byte-identical), so the fixtures pin the GPU
graphdump without a GPU in the dev container.  Run in the dev container only:   python tests/golden/make_gfa_golden.py
The FASTA files are passed by relative name from inside their directory (gfa1 prints the name, UR:Z:<file>)."""
from __future__ import annotations

import hashlib
import json
import os
import subprocess
import sys
import tempfile
from pathlib import Path

ROOT = Path(__file__).resolve().parents[2]
sys.path.insert(0, str(ROOT))

from oracle import oracle as O  # noqa: E402
from tests.cases import GFA_CASES, GFA_FORMATS, build_case  # noqa: E402


def reference_text(d: str, names: list[str], k: int, fmt: str, prefix: bool, image_name: str = "image.dbg") -> bytes:
    cmd = [str(O.REF_GRAPHDUMP), "-f", fmt, "-k", str(k)] + [a for n in names for a in ("-s", n)] + (["--prefix"] if prefix else []) + [image_name]
    q = subprocess.run(cmd, capture_output=True, cwd=d)
    assert q.returncode == 0, q.stderr
    return q.stdout


def main() -> None:
    assert O.REF_GRAPHDUMP.exists(), "build oracle/_ref first: make -C oracle"
    out = {}
    for name, spec in GFA_CASES.items():
        files = build_case(spec)
        with tempfile.TemporaryDirectory() as d:
            recs = []
            for fname, content in files:
                Path(d, fname).write_bytes(content)
                recs += O.parse_fasta(os.path.join(d, fname))
            image, nj, nm = O.find_junctions(recs, spec["k"])
            Path(d, "image.dbg").write_bytes(image)
            ent = {"k": spec["k"], "image_md5": hashlib.md5(image).hexdigest()}
            for fmt, prefix in GFA_FORMATS:
                text = reference_text(d, [f for f, _ in files], spec["k"], fmt, prefix)
                ent[fmt + ("_prefix" if prefix else "")] = {"md5": hashlib.md5(text).hexdigest(), "bytes": len(text)}
            out[name] = ent
            print(name, ent)
    with open(ROOT / "tests" / "golden" / "gfa_golden.json", "w") as fh:
        json.dump(out, fh, indent=1, sort_keys=True)


if __name__ == "__main__":
    main()
