#!/usr/bin/env python
"""Generate tests/golden/golden.json by running the UNMODIFIED reference binary
(oracle/_ref/twopaco, built by oracle/Makefile from /root/reference) on seeded inputs.

Run in the dev container only:   python tests/golden/make_golden.py
The fixtures pin (a) the C oracle and (b) the CUDA path to the reference's own output:
for every case the canonical relabelling of the reference's
de_bruijn.bin is stored as an md5 plus record/junction counts; small edge cases store
the whole canonical stream.  Reference ids are seed-dependent, the canonical stream is not
(verified below by running every case at several -r / -t settings).
"""
from __future__ import annotations

import hashlib
import json
import os
import re
import sys
import tempfile
from pathlib import Path

ROOT = Path(__file__).resolve().parents[2]
sys.path.insert(0, str(ROOT))

from oracle import oracle as O  # noqa: E402
from tests.cases import CASES, build_case  # noqa: E402


def canon_md5(image: bytes) -> str:
    seq, pos, cid = O.canon(image)
    h = hashlib.md5()
    h.update(seq.astype("<i8").tobytes()); h.update(pos.astype("<u4").tobytes()); h.update(cid.astype("<i8").tobytes())
    return h.hexdigest()


def main() -> None:
    assert O.have_reference(), "build oracle/_ref first: make -C oracle"
    # `make_golden.py <case> ...` regenerates only the named cases and keeps the other entries of golden.json
    only = set(sys.argv[1:])
    path = ROOT / "tests" / "golden" / "golden.json"
    out = json.loads(path.read_text()) if only and path.exists() else {}
    for name, spec in CASES.items():
        if only and name not in only:
            continue
        files, k = build_case(spec), spec["k"]
        with tempfile.TemporaryDirectory() as d:
            paths = []
            for i, (fname, content) in enumerate(files):
                p = os.path.join(d, fname)
                with open(p, "wb") as fh:
                    fh.write(content)
                paths.append(p)
            variants = spec.get("ref_variants", [dict(r=1, t=1), dict(r=3, t=4)])
            md5s, logs = set(), []
            for v in variants:
                img, log = O.run_reference(paths, k, spec.get("f", 24), q=spec.get("q", 5), r=v["r"], t=v["t"],
                                           abundance=spec.get("abundance"))
                md5s.add(canon_md5(img)); logs.append(log)
            assert len(md5s) == 1, f"{name}: reference output not canonical-invariant: {md5s}"
            seq, pos, cid = O.canon(img)
            ent = {
                "k": k,
                "input_md5": hashlib.md5(b"\0".join(c for _, c in files)).hexdigest(),
                "canon_md5": md5s.pop(),
                "records": int(len(pos)),
                "image_bytes": len(img),
                "distinct_junctions": int(re.search(r"Distinct junctions = (\d+)", logs[0]).group(1)),
                "true_marks": int(re.search(r"True marks count: (\d+)", logs[0]).group(1)),
            }
            if len(pos) <= 64:
                ent["canon_stream"] = [[int(a), int(b), int(c)] for a, b, c in zip(seq, pos, cid)]
            out[name] = ent
            print(name, ent["records"], ent["distinct_junctions"], ent["canon_md5"])
    with open(path, "w") as fh:
        json.dump(out, fh, indent=1, sort_keys=True)


if __name__ == "__main__":
    main()
