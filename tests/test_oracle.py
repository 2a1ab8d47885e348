"""CPU: pin the C oracle (oracle/junction_oracle.c) to the reference.

* golden.json holds the canonical stream of the UNMODIFIED reference's output per case
  (tests/golden/make_golden.py); the oracle must reproduce every one bit-exactly.
* example.dbg is the reference repository's own shipped golden output (example/README.md:6).
* when oracle/_ref/twopaco is present the reference is also run live on fresh seeds.
"""
import hashlib

import numpy as np
import pytest

from oracle import oracle as O
from tests.cases import CASES, GOLDEN_DIR
from tests.util import canon_md5, case_files, input_md5, oracle_on_paths
from twopaco_b200 import synth


@pytest.mark.parametrize("name", sorted(CASES))
def test_oracle_matches_reference_golden(name, golden):
    spec, g = CASES[name], golden[name]
    with case_files(spec) as (paths, files, _):
        assert input_md5(files) == g["input_md5"], "case generator drifted; regenerate golden.json"
        img, nj, nm = oracle_on_paths(paths, spec["k"], spec.get("abundance", 2**64 - 1))
    assert len(img) == g["image_bytes"]
    assert nj == g["distinct_junctions"]
    assert nm == g["true_marks"] == g["records"]
    assert canon_md5(img) == g["canon_md5"]
    if "canon_stream" in g:
        seq, pos, cid = O.canon(img)
        assert [[int(a), int(b), int(c)] for a, b, c in zip(seq, pos, cid)] == g["canon_stream"]


def test_example_shipped_dbg_and_survey_md5():
    recs = O.parse_fasta(str(GOLDEN_DIR / "example.fa"))
    img, nj, _ = O.find_junctions(recs, 11)
    assert nj == 7 and len(img) == 204
    assert O.canon_equal(img, (GOLDEN_DIR / "example.dbg").read_bytes())
    lst = [(int(a), int(b), int(c)) for a, b, c in zip(*O.canon(img))]
    # SURVEY.md appendix A golden vector
    assert hashlib.md5(repr(lst).encode()).hexdigest() == "82a89c3a7d52fa6a6d294cc55916cf52"


def test_oracle_output_is_already_canonical():
    recs = synth.founder_family(9, 5, 2, 20_000, 0.01, n_runs=1)
    img, nj, _ = O.find_junctions(recs, 25)
    seq, pos, ids = O.decode(img)
    junction = np.abs(ids) <= nj
    # junction ids are numbered by first appearance, first occurrence positive
    first = {}
    nxt = 1
    for i in ids[junction]:
        a = abs(int(i))
        if a not in first:
            assert a == nxt and i > 0
            first[a] = True
            nxt += 1
    # stubs occur exactly once and start at J + 42
    stubs = ids[~junction]
    assert len(set(stubs.tolist())) == len(stubs) and (len(stubs) == 0 or stubs.min() == nj + 42)


def test_invalid_inputs():
    with pytest.raises(O.OracleError):
        O.find_junctions([b"ACGT"], 4)  # even k (constructor.cpp:36-50)
    with pytest.raises(O.OracleError):
        O.find_junctions([b"ACGT"], 641)  # vertexenumerator.cpp:56-70


def test_invalid_fasta_char(tmp_path):
    p = tmp_path / "bad.fa"
    p.write_bytes(b">x\nACGTJACGT\n")
    with pytest.raises(O.OracleError, match="invalid character"):
        O.parse_fasta(str(p))


@pytest.mark.skipif(not O.have_reference(), reason="oracle/_ref not built")
@pytest.mark.parametrize("seed,k,r", [(11, 5, 1), (12, 9, 3), (13, 25, 2), (14, 63, 1)])
def test_oracle_vs_live_reference(seed, k, r, tmp_path):
    recs = synth.reference_selftest_set(seed) if k < 11 else synth.founder_family(seed, 5, 2, 30_000, 0.01, n_runs=2)
    p = tmp_path / "in.fa"
    O.write_fasta(str(p), recs)
    ref_img, _ = O.run_reference([str(p)], k, 22, q=3, r=r, t=2)
    img, _, _ = O.find_junctions(O.parse_fasta(str(p)), k)
    assert O.canon_equal(img, ref_img)


# ---- graphdump -f gfa1 / gfa2 / fasta: pin the Python restatement (oracle/graphdump_gfa.py) to the reference binary's output
@pytest.mark.parametrize("name", ["example_k11", "family_twofiles_k25", "gfa_long_k25", "gfa_mixed_k11", "gfa_mixed_k5"])
def test_gfa_restatement_matches_reference_golden(name, monkeypatch):
    import json
    from oracle import graphdump_gfa as G
    from tests.cases import GFA_CASES, GFA_FORMATS
    spec = GFA_CASES[name]
    g = json.loads((GOLDEN_DIR / "gfa_golden.json").read_text())[name]
    with case_files(spec) as (paths, files, d):
        monkeypatch.chdir(d)                      # gfa1 prints the FASTA file names as they were given
        img, _, _ = oracle_on_paths(paths, spec["k"])
        assert hashlib.md5(img).hexdigest() == g["image_md5"]
        for fmt, prefix in GFA_FORMATS:
            text = G.graphdump_text(img, fmt, spec["k"], [f for f, _ in files], prefix)
            want = g[fmt + ("_prefix" if prefix else "")]
            assert len(text) == want["bytes"] and hashlib.md5(text).hexdigest() == want["md5"], (fmt, prefix)


def test_gfa_restatement_rejects_what_the_reference_rejects(tmp_path, monkeypatch):
    from oracle import graphdump_gfa as G
    from tests.cases import EDGE_LEADING_SHORT
    monkeypatch.chdir(tmp_path)
    (tmp_path / "x.fa").write_bytes(EDGE_LEADING_SHORT)       # sequences shorter than k: no records for them
    img, _, _ = O.find_junctions(O.parse_fasta("x.fa"), 5)
    with pytest.raises(G.GraphdumpError, match="corrupted"):
        G.graphdump_text(img, "gfa1", 5, ["x.fa"])


@pytest.mark.skipif(not O.REF_GRAPHDUMP.exists(), reason="oracle/_ref/graphdump was not built")
@pytest.mark.parametrize("name", ["family_twofiles_k25", "gfa_long_k25", "gfa_mixed_k11"])
def test_gfa_restatement_matches_the_live_reference_on_relabelled_images(name, monkeypatch):
    """Ids and signs as a differently seeded reference run would write them (the fixtures only hold first-appearance numbering):
    the restatement must still print what the unmodified graphdump prints."""
    import subprocess
    from oracle import graphdump_gfa as G
    from tests.cases import GFA_CASES, GFA_FORMATS
    spec = GFA_CASES[name]
    with case_files(spec) as (paths, files, d):
        monkeypatch.chdir(d)
        img, _, _ = oracle_on_paths(paths, spec["k"])
        seq, pos, ids = O.decode(img)
        rng = np.random.default_rng(23)
        uniq = np.unique(np.abs(ids))
        perm = dict(zip(uniq.tolist(), (rng.permutation(len(uniq)) + 3).tolist()))
        flip = {u: int(rng.integers(0, 2)) * 2 - 1 for u in uniq.tolist()}
        rec = np.frombuffer(img, dtype=O.REC_DTYPE).copy()
        sep = (rec["pos"] == O.SEP_POS) | (rec["id"] == O.SEP_ID)
        rec["id"][~sep] = [perm[abs(v)] * (1 if v > 0 else -1) * flip[abs(v)] for v in ids.tolist()]
        with open("relabelled.dbg", "wb") as fh:
            fh.write(rec.tobytes())
        names = [f for f, _ in files]
        for fmt, prefix in GFA_FORMATS:
            cmd = [str(O.REF_GRAPHDUMP), "-f", fmt, "-k", str(spec["k"])] + [a for n in names for a in ("-s", n)] + (["--prefix"] if prefix else [])
            q = subprocess.run(cmd + ["relabelled.dbg"], capture_output=True)
            assert q.returncode == 0, q.stderr
            assert G.graphdump_text(rec.tobytes(), fmt, spec["k"], names, prefix) == q.stdout, (fmt, prefix)
