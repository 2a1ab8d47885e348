"""oracle/oracle.py -- TEST INFRASTRUCTURE ONLY (ctypes front-end of liboracle.so).

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference
legs may import this module.  Nothing under twopaco_b200/ does.

Also hosts the canonical relabelling of a de_bruijn.bin image (SURVEY.md appendix C;
the parity definition of SURVEY.md section 8(c)) and a runner for the unmodified
reference binary built by oracle/Makefile into oracle/_ref/.
"""
from __future__ import annotations

import ctypes
import os
import subprocess
import tempfile
from pathlib import Path

import numpy as np

HERE = Path(__file__).resolve().parent
LIB_PATH = HERE / "liboracle.so"
REF_TWOPACO = HERE / "_ref" / "twopaco"
REF_GRAPHDUMP = HERE / "_ref" / "graphdump"

SEP_POS = 0xFFFFFFFF
SEP_ID = np.iinfo(np.int64).max

_lib = None


def build() -> None:
    """Compile liboracle.so (and oracle/_ref when /root/reference is present)."""
    subprocess.run(["make", "-s", "-C", str(HERE), "all"], check=True)


def lib() -> ctypes.CDLL:
    global _lib
    if _lib is None:
        if not LIB_PATH.exists():
            build()
        L = ctypes.CDLL(str(LIB_PATH))
        L.oracle_last_error.restype = ctypes.c_char_p
        L.oracle_parse_fasta.restype = ctypes.c_int64
        L.oracle_parse_fasta.argtypes = [ctypes.c_char_p, ctypes.POINTER(ctypes.POINTER(ctypes.c_char_p)),
                                         ctypes.POINTER(ctypes.POINTER(ctypes.c_uint64))]
        L.oracle_free_records.argtypes = [ctypes.POINTER(ctypes.c_char_p), ctypes.POINTER(ctypes.c_uint64), ctypes.c_uint64]
        L.oracle_find_junctions.restype = ctypes.c_int
        L.oracle_find_junctions.argtypes = [ctypes.POINTER(ctypes.c_char_p), ctypes.POINTER(ctypes.c_uint64), ctypes.c_uint64,
                                            ctypes.c_uint32, ctypes.c_uint64,
                                            ctypes.POINTER(ctypes.POINTER(ctypes.c_uint8)), ctypes.POINTER(ctypes.c_uint64),
                                            ctypes.POINTER(ctypes.c_uint64), ctypes.POINTER(ctypes.c_uint64)]
        L.oracle_free.argtypes = [ctypes.c_void_p]
        _lib = L
    return _lib


class OracleError(RuntimeError):
    pass


def parse_fasta(path: str) -> list[bytes]:
    """Records of one FASTA file, normalised like the reference (upper-case, non-ACGT -> N)."""
    L = lib()
    seqs = ctypes.POINTER(ctypes.c_char_p)()
    lens = ctypes.POINTER(ctypes.c_uint64)()
    n = L.oracle_parse_fasta(os.fsencode(path), ctypes.byref(seqs), ctypes.byref(lens))
    if n < 0:
        raise OracleError(L.oracle_last_error().decode())
    out = [ctypes.string_at(seqs[i], lens[i]) for i in range(n)]
    L.oracle_free_records(seqs, lens, n)
    return out


def find_junctions(records: list[bytes], k: int, abundance: int = 2**64 - 1) -> tuple[bytes, int, int]:
    """-> (de_bruijn.bin image, distinct junctions, total marks)."""
    L = lib()
    n = len(records)
    arr = (ctypes.c_char_p * max(n, 1))(*records)
    lens = (ctypes.c_uint64 * max(n, 1))(*[len(r) for r in records])
    out = ctypes.POINTER(ctypes.c_uint8)()
    nbytes = ctypes.c_uint64()
    nj = ctypes.c_uint64()
    nm = ctypes.c_uint64()
    rc = L.oracle_find_junctions(arr, lens, n, k, abundance, ctypes.byref(out), ctypes.byref(nbytes),
                                 ctypes.byref(nj), ctypes.byref(nm))
    if rc != 0:
        raise OracleError(L.oracle_last_error().decode())
    data = ctypes.string_at(out, nbytes.value) if nbytes.value else b""
    L.oracle_free(out)
    return data, nj.value, nm.value


REC_DTYPE = np.dtype([("pos", "<u4"), ("id", "<i8")])  # 12 bytes, junctionapi.h:125-126


def decode(image: bytes):
    """de_bruijn.bin image -> (seq[int64], pos[uint32], id[int64]) without separators
    (reader semantics: junctionapi.h:81-99)."""
    if len(image) % 12:
        raise ValueError("image size is not a multiple of 12")
    rec = np.frombuffer(image, dtype=REC_DTYPE)
    sep = (rec["pos"] == SEP_POS) | (rec["id"] == SEP_ID)
    seq = np.cumsum(sep)[~sep].astype(np.int64)
    return seq, rec["pos"][~sep].copy(), rec["id"][~sep].copy()


def canon(image: bytes):
    """Canonical relabelling: ids renumbered by first appearance,
    first occurrence positive.  Two images are equivalent iff their canon() are equal."""
    seq, pos, ids = decode(image)
    a = np.abs(ids)
    uniq, first_idx, inv = np.unique(a, return_index=True, return_inverse=True)
    order = np.argsort(first_idx, kind="stable")          # uniq index -> rank of first appearance
    new_of_uniq = np.empty(len(uniq), dtype=np.int64)
    new_of_uniq[order] = np.arange(1, len(uniq) + 1)
    first_sign = np.sign(ids[first_idx])                   # sign at first appearance, per uniq
    cid = new_of_uniq[inv] * np.sign(ids) * first_sign[inv]
    return seq, pos, cid.astype(np.int64)


def canon_equal(a: bytes, b: bytes) -> bool:
    ca, cb = canon(a), canon(b)
    return all(np.array_equal(x, y) for x, y in zip(ca, cb))


def write_fasta(path: str, records: list[bytes], width: int = 80, names: list[str] | None = None) -> None:
    with open(path, "wb") as f:
        for i, r in enumerate(records):
            f.write(b">" + (names[i].encode() if names else b"s%d" % i) + b"\n")
            for j in range(0, len(r), width):
                f.write(r[j:j + width] + b"\n")
    # reference quirk: sizes that are exact multiples of 2^20 trip a stale-buffer read
    if os.path.getsize(path) % (1 << 20) == 0:
        with open(path, "ab") as f:
            f.write(b"\n")


def have_reference() -> bool:
    return REF_TWOPACO.exists() and os.access(REF_TWOPACO, os.X_OK)


def run_reference(fasta: list[str], k: int, f: int, q: int = 5, r: int = 1, t: int = 1,
                  abundance: int | None = None, timeout: float | None = None) -> tuple[bytes, str]:
    """Run the unmodified reference twopaco (oracle/_ref) -> (de_bruijn.bin image, log)."""
    with tempfile.TemporaryDirectory(prefix="tpc_ref_") as d:
        out = os.path.join(d, "out.bin")
        cmd = [str(REF_TWOPACO), "-k", str(k), "-f", str(f), "-q", str(q), "-r", str(r), "-t", str(t),
               "--tmpdir", d, "-o", out]
        if abundance is not None:
            cmd += ["-a", str(abundance)]
        cmd += list(fasta)
        p = subprocess.run(cmd, capture_output=True, text=True, timeout=timeout)
        if p.returncode != 0:
            raise OracleError(f"reference failed ({p.returncode}): {p.stderr.strip()}")
        with open(out, "rb") as fh:
            return fh.read(), p.stdout
