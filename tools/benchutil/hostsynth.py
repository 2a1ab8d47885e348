"""numpy restatement of the on-device founder-family generator (tools/benchutil/tpc_benchutil.cu:
founder_base / synth_emit) -- BENCH / TEST INFRASTRUCTURE ONLY.  Lets `bench.py --impl reference` build its
bounded sample of a workload on host cores, without a GPU and without loading the product library."""
from __future__ import annotations

import numpy as np

M64 = np.uint64(0xFFFFFFFFFFFFFFFF)


def _fmix64(x: np.ndarray) -> np.ndarray:
    with np.errstate(over="ignore"):
        x = x ^ (x >> np.uint64(33)); x = x * np.uint64(0xff51afd7ed558ccd)
        x = x ^ (x >> np.uint64(33)); x = x * np.uint64(0xc4ceb9fe1a85ec53)
        return x ^ (x >> np.uint64(33))


def record_prefix(seed: int, records_per_genome: int, p: float, g: int, c: int, nbases: int, record_len: int) -> bytes:
    """The first `nbases` bases of record c of genome g (genome 0 = founder) of the family generated from `seed`."""
    want = min(nbases, record_len)
    n = min(record_len, int(want * 1.02) + 64)              # founder positions to expand (deletions shorten the output)
    i = np.arange(n, dtype=np.uint64)
    with np.errstate(over="ignore"):
        fb = ((_fmix64((np.uint64(seed) ^ np.uint64(0xF00DFACE5EED)) + ((np.uint64(c) << np.uint64(36)) | i)) >> np.uint64(17))
              & np.uint64(3)).astype(np.uint8)
        if g == 0:
            out = fb
        else:
            r = _fmix64(np.uint64(seed) + _fmix64((np.uint64(g * records_per_genome + c) << np.uint64(36)) | i))
            thr = np.uint64(int(p * 4294967296.0))
            mut = (r & np.uint64(0xFFFFFFFF)) < thr
            kind = (r >> np.uint64(32)) & np.uint64(0xFFFF)
            extra = (r >> np.uint64(48)).astype(np.uint32)
            snp = mut & (kind < np.uint64(52429))
            ins = mut & ~snp & (kind < np.uint64(58982))
            dele = mut & ~snp & ~ins
            b0 = np.where(snp, (fb.astype(np.uint32) + 1 + extra % 3) & 3, fb).astype(np.uint8)
            b1 = (extra & 3).astype(np.uint8)
            count = np.where(dele, 0, np.where(ins, 2, 1))
            start = np.cumsum(count) - count
            out = np.empty(int(count.sum()), dtype=np.uint8)
            keep = ~dele
            out[start[keep]] = b0[keep]
            out[start[ins] + 1] = b1[ins]
    assert len(out) >= want or n == record_len, "expansion margin too small"
    return np.frombuffer(b"ACGT", dtype=np.uint8)[out[:want]].tobytes()
