"""Helpers shared by the CPU and GPU parity tests."""
from __future__ import annotations

import hashlib
import os
import tempfile
from contextlib import contextmanager

from oracle import oracle as O
from tests.cases import build_case


def canon_md5(image: bytes) -> str:
    seq, pos, cid = O.canon(image)
    h = hashlib.md5()
    h.update(seq.astype("<i8").tobytes()); h.update(pos.astype("<u4").tobytes()); h.update(cid.astype("<i8").tobytes())
    return h.hexdigest()


def input_md5(files) -> str:
    return hashlib.md5(b"\0".join(c for _, c in files)).hexdigest()


@contextmanager
def case_files(spec):
    """Materialise a case's FASTA files in a temp dir -> list of paths."""
    files = build_case(spec)
    with tempfile.TemporaryDirectory(prefix="tpc_case_") as d:
        paths = []
        for name, content in files:
            p = os.path.join(d, name)
            with open(p, "wb") as fh:
                fh.write(content)
            paths.append(p)
        yield paths, files, d


def oracle_on_paths(paths, k, abundance=2**64 - 1):
    recs = []
    for p in paths:
        recs += O.parse_fasta(p)
    return O.find_junctions(recs, k, abundance)
