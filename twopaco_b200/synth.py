"""Seeded synthetic genome sets (host side, numpy) -- the input shapes of SURVEY.md section 8(d).

* ``founder_family``: star phylogeny used by the BASELINE configs (founder i.i.d. over ACGT;
  each derived genome mutates every founder base with probability p: 80 % SNP, 10 % 1-bp
  insertion, 10 % 1-bp deletion).
* ``reference_selftest_set``: the recipe of the reference's own ``--test`` mode
  (/root/reference/src/graphconstructor/test.cpp:20-67 and constructor.cpp:147): one random
  chromosome with 'N' at rate 1/500 plus mutated copies (change rate 0.05; of the changes 10 %
  substitution by a random base, 45 % insertion after the base, 45 % deletion).

Everything is deterministic in ``seed``.  Records are ``bytes`` over ACGTN.
"""
from __future__ import annotations

import numpy as np

ALPHABET = np.frombuffer(b"ACGT", dtype=np.uint8)


def random_dna(rng: np.random.Generator, n: int) -> np.ndarray:
    return ALPHABET[rng.integers(0, 4, size=n, dtype=np.uint8)]


def mutate(rng: np.random.Generator, base: np.ndarray, p: float, snp: float = 0.8, ins: float = 0.1) -> np.ndarray:
    """SNP / 1-bp insertion / 1-bp deletion at rate p per founder base."""
    n = len(base)
    hit = rng.random(n) < p
    kind = rng.random(n)
    out = base.copy()
    is_snp = hit & (kind < snp)
    # uniform *other* base
    idx = np.searchsorted(ALPHABET, base[is_snp])
    out[is_snp] = ALPHABET[(idx + rng.integers(1, 4, size=idx.shape[0])) % 4]
    is_ins = hit & (kind >= snp) & (kind < snp + ins)
    is_del = hit & (kind >= snp + ins)
    reps = np.ones(n, dtype=np.int64)
    reps[is_del] = 0
    reps[is_ins] = 2
    res = np.repeat(out, reps)
    # the second copy of an inserted base becomes a fresh uniform base
    ins_pos = np.cumsum(reps)[is_ins] - 1
    res[ins_pos] = ALPHABET[rng.integers(0, 4, size=ins_pos.shape[0])]
    return res


def founder_family(seed: int, genomes: int, records_per_genome: int, record_len: int, p: float,
                   n_runs: int = 0, n_run_len: int = 50) -> list[bytes]:
    """genomes x records_per_genome records; genome 0 is the founder.  Optional runs of 'N'."""
    rng = np.random.default_rng(seed)
    founders = [random_dna(rng, record_len) for _ in range(records_per_genome)]
    out: list[bytes] = []
    for g in range(genomes):
        for c in range(records_per_genome):
            s = founders[c] if g == 0 else mutate(rng, founders[c], p)
            s = s.copy()
            for _ in range(n_runs):
                a = int(rng.integers(0, max(1, len(s) - n_run_len)))
                s[a:a + int(rng.integers(1, n_run_len + 1))] = ord("N")
            out.append(s.tobytes())
    return out


def reference_selftest_set(seed: int, length: int = 9000, copies: int = 6, change: float = 0.05,
                           subst: float = 0.1, n_rate: float = 1.0 / 500) -> list[bytes]:
    rng = np.random.default_rng(seed)
    chr0 = random_dna(rng, length)
    chr0[rng.random(length) < n_rate] = ord("N")
    out = [chr0.tobytes()]
    for _ in range(1, copies):
        res = bytearray()
        ch_hit = rng.random(length) <= change
        ch_sub = rng.random(length) <= subst
        ch_ins = rng.random(length) <= 0.5
        rnd = ALPHABET[rng.integers(0, 4, size=length)]
        for i in range(length):
            if ch_hit[i]:
                if ch_sub[i]:
                    res.append(rnd[i])
                elif ch_ins[i]:
                    res.append(chr0[i]); res.append(rnd[i])
            else:
                res.append(chr0[i])
        out.append(bytes(res))
    return out
