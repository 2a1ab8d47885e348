"""CPU: the C-ABI library loads and exports every symbol include/twopaco_b200.h declares; the
host-side pieces (FASTA framing, 2-bit packing) match the oracle / a numpy restatement.
No GPU compute is called here."""
import os
import re
from pathlib import Path

import numpy as np
import pytest

from oracle import oracle as O
from tests.cases import CASES
from tests.util import case_files
from twopaco_b200 import api, synth

ROOT = Path(__file__).resolve().parents[1]


def test_every_declared_symbol_is_exported_and_bound():
    header = (ROOT / "include" / "twopaco_b200.h").read_text()
    header = re.sub(r"/\*.*?\*/", "", header, flags=re.S)  # declarations only
    declared = set(re.findall(r"\b(tpc_[a-z0-9_]+)\s*\(", header)) - {"tpc_log_fn"}
    L = api.lib()
    for name in sorted(declared):
        assert hasattr(L, name), f"{name} declared in the header but not exported"
    assert declared == set(api.SIGNATURES), declared ^ set(api.SIGNATURES)
    assert L.tpc_abi_version() == 4


def numpy_pack(records):
    """Independent restatement of the layout in include/twopaco_b200.h (tpc_genome)."""
    npos = 1 + sum(len(r) + 1 for r in records)
    code = np.zeros(npos, dtype=np.uint64)
    isn = np.ones(npos, dtype=np.uint64)
    lut = np.full(256, 4, dtype=np.uint8)
    for i, ch in enumerate(b"ACGT"):
        lut[ch] = i
        lut[ord(chr(ch).lower())] = i
    p = 1
    starts = []
    for r in records:
        starts.append(p)
        c = lut[np.frombuffer(r, dtype=np.uint8)]
        code[p:p + len(r)] = np.where(c < 4, c, 0)
        isn[p:p + len(r)] = (c >= 4)
        p += len(r) + 1
    return npos, code, isn, starts


@pytest.mark.parametrize("threads", [1, 3])
def test_pack_records_layout(threads):
    recs = synth.founder_family(5, 3, 2, 10_000, 0.01, n_runs=2) + [b"", b"ACG", b"acgtnNRY" * 3]
    g = api.pack_records(recs, threads=threads)
    npos, code, isn, starts = numpy_pack(recs)
    assert g.n_positions == npos and list(g.rec_start) == starts
    L = api.lib()
    assert len(g.codes) == L.tpc_code_words(npos) and len(g.n_mask) == L.tpc_mask_words(npos)
    pos = np.arange(npos, dtype=np.uint64)
    got_code = (g.codes[pos >> np.uint64(5)] >> (np.uint64(2) * (pos & np.uint64(31)))) & np.uint64(3)
    got_n = (g.n_mask[pos >> np.uint64(6)] >> (pos & np.uint64(63))) & np.uint64(1)
    assert np.array_equal(got_code, code) and np.array_equal(got_n, isn)
    # padding: codes 0, n_mask 1 beyond n_positions
    tail = np.arange(npos, len(g.n_mask) * 64, dtype=np.uint64)
    assert np.all((g.n_mask[tail >> np.uint64(6)] >> (tail & np.uint64(63))) & np.uint64(1) == 1)


@pytest.mark.parametrize("name", ["example_k11", "edge_mixed_k11", "edge_leading_short_k5", "family_twofiles_k25"])
def test_read_fasta_matches_oracle(name):
    with case_files(CASES[name]) as (paths, _, _):
        ours = api.read_fasta(paths)
        ref = []
        for p in paths:
            ref += O.parse_fasta(p)
    assert ours == ref


def test_read_fasta_errors(tmp_path):
    with pytest.raises(api.TpcError, match="Can't open file"):
        api.read_fasta([str(tmp_path / "missing.fa")])
    bad = tmp_path / "bad.fa"
    bad.write_bytes(b">x y\nACGTJACGT\n")
    with pytest.raises(api.TpcError, match="invalid character 'J' in sequence x"):
        api.read_fasta([str(bad)])
    bad.write_bytes(b"ACGT\n")
    with pytest.raises(api.TpcError, match="should start with a '>'"):
        api.read_fasta([str(bad)])


def test_parameter_validation_and_no_cpu_fallback():
    import torch
    with pytest.raises(api.TpcError, match="must be odd"):
        api.Session(k=24, filter_bits=20)
    with pytest.raises(api.TpcError, match="K is too big"):
        api.Session(k=605, filter_bits=20)   # reference: capacity ceil((k + 4) / 32) reaches MAX_CAPACITY 20
    if not torch.cuda.is_available():
        # the product path must fail loudly without a GPU
        with pytest.raises(api.TpcError, match="no CUDA device"):
            api.Session(k=25, filter_bits=20)


@pytest.mark.parametrize("threads,piece", [(1, 0), (4, 0), (3, 7), (4, 1000)])
@pytest.mark.parametrize("name", ["example_k11", "edge_mixed_k11", "edge_leading_short_k5", "family_twofiles_k25", "family_seam_k25"])
def test_multithreaded_ingest_matches_oracle_parser(name, threads, piece, monkeypatch):
    """tpc_ingest_fasta (what tpc_build uses) == the oracle's restatement of StreamFastaParser.
    Tiny work units (TPC_INGEST_PIECE) exercise the seams between pieces and between staging spans."""
    if piece:
        if name == "family_seam_k25" and piece < 100:
            pytest.skip("too slow with 7-byte pieces")
        monkeypatch.setenv("TPC_INGEST_PIECE", str(piece))
    with case_files(CASES[name]) as (paths, _, _):
        layout, npos, rec_start, rec_len = api.ingest_fasta(paths, threads=threads)
        ref = []
        for p in paths:
            ref += O.parse_fasta(p)
    assert [int(x) for x in rec_len] == [len(r) for r in ref]
    expect = b"N" + b"".join(r + b"N" for r in ref)
    assert npos == len(expect) and layout == expect
    assert [int(x) for x in rec_start] == list(np.cumsum([1] + [len(r) + 1 for r in ref])[:-1])


def test_ingest_errors_and_odd_framing(tmp_path):
    with pytest.raises(api.TpcError, match="Can't open file"):
        api.ingest_fasta([str(tmp_path / "missing.fa")])
    f = tmp_path / "x.fa"
    f.write_bytes(b">x y\nACGTJACGT\n")
    with pytest.raises(api.TpcError, match="invalid character 'J' in sequence x"):
        api.ingest_fasta([str(f)])
    f.write_bytes(b"ACGT\n")
    with pytest.raises(api.TpcError, match="should start with a '>'"):
        api.ingest_fasta([str(f)])
    # '>' inside a header line is header text; '>' inside sequence text starts a record; a header
    # without newline is an empty record; empty files have no records (streamfastaparser.cpp:29-93)
    f.write_bytes(b">a >not a record\nAC\r\nGT>b\nTT\n\n>c")
    layout, npos, rs, rl = api.ingest_fasta([str(f)], threads=3)
    assert layout == b"NACGTNTTNN" and list(rl) == [4, 2, 0]
    assert O.parse_fasta(str(f)) == [b"ACGT", b"TT", b""]
    e = tmp_path / "empty.fa"
    e.write_bytes(b"")
    layout, npos, rs, rl = api.ingest_fasta([str(e), str(f)])
    assert layout == b"NACGTNTTNN"


@pytest.mark.parametrize("seed", range(6))
def test_ingest_fuzz_against_the_scalar_parsers(seed, tmp_path, monkeypatch):
    """The 32-byte-at-a-time (AVX2) classification of tpc_ingest_fasta against tpc_read_fasta (scalar,
    streamfastaparser.cpp:29-133) and the oracle's parser on random files: every letter of the alphabet in both
    cases, all whitespace kinds anywhere in the sequence text, ragged lines, records shorter than a SIMD block,
    '>' inside headers, no trailing newline; pieces of odd sizes so that blocks straddle piece seams."""
    import random
    rnd = random.Random(seed)
    alphabet = "ACGTURYKMSWBDHWNXV"
    letters = alphabet + alphabet.lower() + "ACGTacgt" * 6
    blanks = [" ", "\t", "\n", "\r\n", "\v", "\f", "\n\n"]
    text = []
    for r in range(rnd.randint(1, 7)):
        text.append(">r%d some >text\t here" % r + rnd.choice(["\n", "\r\n"]))
        for _ in range(rnd.choice([0, 1, 3, 40, 400])):
            n = rnd.choice([0, 1, 5, 31, 32, 33, 63, 64, 80, 80, 80, 200])
            text.append("".join(rnd.choice(letters) for _ in range(n)))
            text.append(rnd.choice(blanks) if rnd.random() < 0.9 else "")
    body = "".join(text)
    if rnd.random() < 0.5:
        body = body.rstrip()
    f = tmp_path / "fuzz.fa"
    f.write_bytes(body.encode())
    ref = O.parse_fasta(str(f))
    assert api.read_fasta([str(f)]) == ref
    for piece in (0, 37, 64, 1001):
        if piece:
            monkeypatch.setenv("TPC_INGEST_PIECE", str(piece))
        layout, npos, rec_start, rec_len = api.ingest_fasta([str(f)], threads=3)
        assert layout == b"N" + b"".join(x + b"N" for x in ref), (seed, piece)
    # an invalid character inside a block, between valid ones
    bad = body.replace("\n", "\n", 1).encode() + b"\n>z\n" + b"ACGT" * 20 + b"!" + b"ACGT" * 20 + b"\n"
    f.write_bytes(bad)
    with pytest.raises(api.TpcError, match="invalid character '!' in sequence z"):
        api.ingest_fasta([str(f)], threads=2)


def test_graphdump_cli_argument_errors(tmp_path):
    """`graphdump` (host/graphdump.cpp) keeps the reference's flags and exit codes (graphdump.cpp:608-710); argument errors
    are reported before any GPU work, so they are checked here."""
    import subprocess
    cli = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "twopaco_b200", "bin", "graphdump")
    if not os.path.exists(cli):
        pytest.skip("CLI not built")
    run = lambda *a: subprocess.run([cli, *a], capture_output=True, text=True, cwd=tmp_path)
    for args, text in ((( ), "infile"), (("-f", "gfa1", "-k", "11", "x.dbg"), "seqfilename"), (("-f", "nope", "-k", "11", "x.dbg"), "does not meet constraint"),
                       (("-f", "seq", "x.dbg"), "kvalue"), (("-k", "11", "x.dbg"), "format"), (("-f", "seq", "-k", "11", "a", "b"), "Too many arguments"),
                       (("-f", "seq", "-k", "eleven", "x.dbg"), "Couldn't read argument value"), (("-f", "seq", "-k", "11", "missing.dbg"), "Can't open file"),
                       (("--bogus", "x.dbg"), "Couldn't find match")):
        r = run(*args)
        assert r.returncode == 1 and r.stderr.startswith("error: ") and text in r.stderr, (args, r.stderr)
    r = run("--help")
    assert r.returncode == 0 and "--prefix" in r.stdout
