"""GPU (-m gpu): the CUDA path through the C ABI, bit-exact against
  * golden.json  (canonical streams of the UNMODIFIED reference's outputs, committed), and
  * the C oracle on the same seeded inputs.
Integer / byte work: the bar is bit-exact equality of the canonical stream."""
import os

import numpy as np
import pytest

from oracle import oracle as O
from tests.cases import CASES
from tests.util import canon_md5, case_files, oracle_on_paths
from tools import benchutil
from twopaco_b200 import api, synth

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("name", sorted(CASES))
def test_create_enumerator_matches_reference_golden(name, golden, tmp_path):
    """Level 1 (tpc_build == CreateEnumerator): FASTA files -> de_bruijn.bin."""
    spec, g = CASES[name], golden[name]
    out = str(tmp_path / "de_bruijn.bin")
    with case_files(spec) as (paths, _, d):
        ve = api.CreateEnumerator(paths, spec["k"], spec.get("f", 24), hashFunctions=spec.get("q", 5), rounds=1, threads=4,
                                  abundance=spec.get("abundance", api.ABUNDANCE_MAX), tmpDirName=d, outFileName=out)
        oracle_img, nj, nm = oracle_on_paths(paths, spec["k"], spec.get("abundance", 2**64 - 1))
    img = open(out, "rb").read()
    assert len(img) == g["image_bytes"]
    assert ve.GetVerticesCount() == g["distinct_junctions"] == nj
    assert canon_md5(img) == g["canon_md5"], "canonical stream differs from the reference's"
    assert O.canon_equal(img, oracle_img)
    st = ve.stats()
    assert st.occurrences == g["true_marks"] == nm
    assert f"True junctions count = {nj}" in ve.log and f"True marks count: {nm}" in ve.log
    # ids: junctions 1..J by first appearance (first occurrence positive), stubs J+42.. in stream order
    seq, pos, ids = O.decode(img)
    _, _, oids = O.decode(oracle_img)
    assert np.array_equal(ids, oids), "ids differ from the oracle's first-appearance numbering"
    ve.close()


@pytest.mark.parametrize("rounds", [2, 3, 4])
@pytest.mark.parametrize("name", ["selftest_s1_k7", "selftest_s2_k9", "family_k25", "family_k63", "family_k161", "edge_mixed_k5"])
def test_rounds_do_not_change_the_result(name, rounds, golden, tmp_path):
    """-r: hash-range rounds (vertexenumerator.h:228-392) -- the reference's --test sweeps 1..4."""
    spec, g = CASES[name], golden[name]
    out = str(tmp_path / "o.bin")
    with case_files(spec) as (paths, _, d):
        ve = api.CreateEnumerator(paths, spec["k"], spec.get("f", 24), hashFunctions=spec.get("q", 5), rounds=rounds,
                                  threads=2, tmpDirName=d, outFileName=out)
    assert canon_md5(open(out, "rb").read()) == g["canon_md5"]
    assert ve.GetVerticesCount() == g["distinct_junctions"]
    ve.close()


@pytest.mark.parametrize("q,f", [(1, 20), (2, 12), (3, 16), (8, 22), (5, 9)])
def test_filter_shape_never_changes_the_result(q, f, golden):
    """Tiny / saturated filters only add false candidates; pass 2 must remove them all."""
    spec, g = CASES["family_k25"], golden["family_k25"]
    with case_files(spec) as (paths, _, _):
        recs = api.read_fasta(paths)
    img, st = api.junctions_host(api.pack_records(recs), k=25, filter_bits=f, q=q)
    assert canon_md5(bytes(img)) == g["canon_md5"]
    assert st.junctions == g["distinct_junctions"] and st.candidate_kmers >= st.junctions


@pytest.mark.parametrize("k", [9, 161])
def test_get_id_surface(k, golden):
    """VertexEnumerator::GetId (vertexenumerator.h:98-102; test.cpp:234-242): every junction
    k-mer has an id, its reverse complement the negated id, anything else INVALID_VERTEX."""
    recs = synth.reference_selftest_set(77) if k == 9 else synth.founder_family(78, 4, 1, 30_000, 0.002, n_runs=1)
    oracle_img, nj, _ = O.find_junctions(recs, k)
    s = api.Session(k=k, filter_bits=20)
    s.set_genome_host(api.pack_records(recs))
    s.find_candidates()
    ptr, n = s.local_junctions()
    assert n == nj
    s.set_junctions(ptr, n)
    seq, pos, ids = O.decode(oracle_img)
    comp = bytes.maketrans(b"ACGT", b"TGCA")
    seen = 0
    for sq, p, i in list(zip(seq, pos, ids))[:400]:
        kmer = recs[sq][p:p + k]
        if abs(i) <= nj:
            assert s.get_id(kmer.decode()) == i
            assert s.get_id(kmer.translate(comp)[::-1].decode()) == -i
            seen += 1
        elif b"N" not in kmer:
            assert s.get_id(kmer.decode()) == api.INVALID_VERTEX
    assert seen > 50
    assert s.get_id("ACGT") == api.INVALID_VERTEX          # wrong length
    assert s.get_id("ACGTNACGT" + "A" * (k - 9)) == api.INVALID_VERTEX     # not definite
    s.close()


def test_empty_and_degenerate_inputs():
    for recs in ([], [b""], [b"ACG"], [b"N" * 40], [b"ACGTACGTACG"], [b"", b"ACGTTGCAAGC", b""]):
        ref, nj, nm = O.find_junctions(recs, 11)
        img, st = api.junctions_host(api.pack_records(recs), k=11, filter_bits=16)
        assert bytes(img) == ref, recs
        assert st.junctions == nj and st.occurrences == nm


def test_sharded_sessions_union_equals_unsharded():
    """Hash-range shards (spatial -r): the union of the shards' junction sets and the OR of
    their masks reproduce the single-GPU result (all shards run on this one GPU here)."""
    import torch
    recs = synth.founder_family(31, 6, 2, 40_000, 0.01, n_runs=2)
    k = 25
    ref, nj, _ = O.find_junctions(recs, k)
    g = api.pack_records(recs)
    shards = [api.Session(k=k, filter_bits=22, shard_index=i, shard_count=3) for i in range(3)]
    lists, masks = [], []
    for s in shards:
        s.set_genome_host(g)
        s.find_candidates()
        ptr, n = s.local_junctions()
        lists.append(_dev_to_torch(ptr, n, torch.int64).clone())
        mptr, mw = s.candidate_mask()
        masks.append(_dev_to_torch(mptr, mw, torch.int32))
    allj = torch.cat(lists)
    assert allj.numel() == nj
    total = masks[0].clone()
    for m in masks[1:]:
        assert int((total & m).ne(0).sum()) == 0, "shard masks must be disjoint"
        total |= m
    s0 = shards[0]
    masks[0].copy_(total)
    s0.set_junctions(allj.data_ptr(), allj.numel())
    nrec, nstub = s0.emit_count(0, g.n_positions)
    out = torch.empty(12 * (nrec + len(recs)) + 16, dtype=torch.uint8, device="cuda")
    off, nb = s0.emit_write(0, 0, out.data_ptr(), out.numel())
    torch.cuda.synchronize()
    assert off == 0 and bytes(out[:nb].cpu().numpy()) == ref
    for s in shards:
        s.close()


def test_position_sliced_emit_concatenates():
    """Position-sharded emit (multi-GPU output path): slices concatenate to the full image."""
    import torch
    recs = synth.founder_family(32, 5, 3, 30_000, 0.01, n_runs=1) + [b"ACG", b""] + synth.founder_family(33, 2, 1, 9_000, 0.02)
    k = 31
    ref, nj, _ = O.find_junctions(recs, k)
    g = api.pack_records(recs)
    s = api.Session(k=k, filter_bits=22)
    s.set_genome_host(g)
    s.find_candidates()
    ptr, n = s.local_junctions()
    s.set_junctions(ptr, n)
    cuts = [0, 8192 * 3, 8192 * 20, 8192 * 21, 8192 * 40, g.n_positions]
    counts = []
    for a, b in zip(cuts, cuts[1:]):
        counts.append(s.emit_count(a, b))
    image = bytearray()
    rb = sb = 0
    for (a, b), (nr, ns) in zip(zip(cuts, cuts[1:]), counts):
        s.emit_count(a, b)
        out = torch.empty(12 * (nr + len(recs)) + 16, dtype=torch.uint8, device="cuda")
        off, nb = s.emit_write(rb, sb, out.data_ptr(), out.numel())
        torch.cuda.synchronize()
        assert off == len(image)
        image += bytes(out[:nb].cpu().numpy())
        rb += nr
        sb += ns
    assert bytes(image) == ref
    s.close()


class _DevArray:
    def __init__(self, ptr, n, typestr):
        self.__cuda_array_interface__ = {"shape": (n,), "typestr": typestr, "data": (ptr, False), "version": 2}


def _dev_to_torch(ptr, n, dtype):
    import torch
    typestr = {torch.int64: "<i8", torch.int32: "<i4", torch.uint8: "|u1"}[dtype]
    if n == 0:
        return torch.empty(0, dtype=dtype, device="cuda")
    return torch.as_tensor(_DevArray(ptr, n, typestr), device="cuda")


@pytest.mark.skipif(not O.have_reference(), reason="oracle/_ref binaries did not travel")
def test_live_reference_binary_on_the_gpu_box(tmp_path):
    recs = synth.founder_family(808, 6, 2, 80_000, 0.01, n_runs=2)
    p = tmp_path / "in.fa"
    O.write_fasta(str(p), recs)
    ref_img, _ = O.run_reference([str(p)], 25, 24, q=5, r=1, t=os.cpu_count() or 1)
    out = str(tmp_path / "o.bin")
    ve = api.CreateEnumerator([str(p)], 25, 24, outFileName=out, tmpDirName=str(tmp_path))
    assert O.canon_equal(open(out, "rb").read(), ref_img)
    ve.close()


def test_full_size_properties_c2_slice():
    """Size-independent properties at a larger size (oracle too slow): idempotence across
    filter shapes / rounds, strand symmetry (reverse-complemented input gives the mirrored
    stream), sortedness, and the id/stub invariants."""
    recs = synth.founder_family(0xEC01, 12, 1, 1_000_000, 0.01)
    g = api.pack_records(recs)
    img_a, st_a = api.junctions_host(g, k=25, filter_bits=30, q=5)
    img_b, st_b = api.junctions_host(g, k=25, filter_bits=24, q=2, rounds=3)
    assert bytes(img_a) == bytes(img_b) and st_a.junctions == st_b.junctions
    seq, pos, ids = O.decode(bytes(img_a))
    order = np.lexsort((pos, seq))
    assert np.array_equal(order, np.arange(len(pos))), "records must be sorted by (seq, pos)"
    J = st_a.junctions
    junction = np.abs(ids) <= J
    assert set(np.abs(ids[junction]).tolist()) == set(range(1, J + 1))
    stubs = ids[~junction]
    assert np.array_equal(stubs, J + 42 + np.arange(len(stubs)))
    # strand symmetry: reverse-complement every record
    comp = bytes.maketrans(b"ACGT", b"TGCA")
    rc = [r.translate(comp)[::-1] for r in recs]
    img_c, st_c = api.junctions_host(api.pack_records(rc), k=25, filter_bits=30)
    assert st_c.junctions == J
    seq_c, pos_c, ids_c = O.decode(bytes(img_c))
    lens = np.array([len(r) for r in recs], dtype=np.int64)
    mirrored = sorted(zip(seq_c.tolist(), (lens[seq_c] - 25 - pos_c.astype(np.int64)).tolist()))
    assert mirrored == sorted(zip(seq.tolist(), pos.astype(np.int64).tolist()))


def test_k0_pack_ascii_device_matches_host_packer():
    """K0 (ASCII -> 2-bit + N mask on the device) == tpc_pack_records on the host."""
    recs = synth.founder_family(41, 3, 2, 70_001, 0.01, n_runs=3) + [b"", b"acgtnRYKM" * 7, b"ACG"]
    host = api.pack_records(recs)
    layout = bytearray(b"N")
    for r in recs:
        layout += r + b"N"
    assert len(layout) == host.n_positions
    buf = api.DeviceBuffer((len(layout) + 63) // 64 * 64 + 64)
    buf.from_host(np.frombuffer(bytes(layout), dtype=np.uint8))
    dg = api.pack_ascii_device(buf, host.n_positions, host.rec_start, host.rec_len)
    dev = dg.to_host()
    assert np.array_equal(dev.codes, host.codes) and np.array_equal(dev.n_mask, host.n_mask)


def test_synth_family_device_generator():
    """On-device founder-family generator: deterministic, right shape, right divergence."""
    a = benchutil.synth_family_device(0x4855, 4, 3, 200_000, 0.01)
    b = benchutil.synth_family_device(0x4855, 4, 3, 200_000, 0.01)
    assert a.n_positions == b.n_positions and np.array_equal(a.codes.to_host(), b.codes.to_host())
    assert len(a.rec_len) == 12 and np.all(a.rec_len[:3] == 200_000)
    assert int(a.rec_start[0]) == 1 and a.n_positions == 1 + int((a.rec_len + 1).sum())
    founder, derived = a.record_ascii(1), a.record_ascii(4)   # record 1 of genome 0 and of genome 1
    assert set(founder) <= set(b"ACGT") and abs(len(derived) - len(founder)) < 200
    # ~1 % divergence: (1 - 0.01)^16 = 85 % of the founder's 16-mers survive in the derived record
    fk = {founder[i:i + 16] for i in range(0, 50_000)}
    dk = {derived[i:i + 16] for i in range(0, 51_000)}
    assert 0.75 < len(fk & dk) / len(fk) < 0.93
    assert founder != a.record_ascii(0) and derived != a.record_ascii(7)
    # the packed form spells the same bases as the ASCII buffer
    host = a.to_host()
    pos = np.arange(int(a.rec_start[4]), int(a.rec_start[4]) + 1000, dtype=np.uint64)
    code = (host.codes[pos >> np.uint64(5)] >> (np.uint64(2) * (pos & np.uint64(31)))) & np.uint64(3)
    assert bytes(np.frombuffer(b"ACGT", np.uint8)[code.astype(np.int64)]) == derived[:1000]
    # and the junction finder agrees with the oracle on it
    recs = [a.record_ascii(r) for r in range(12)]
    ref, nj, _ = O.find_junctions(recs, 25)
    s = api.Session(k=25, filter_bits=26)
    a.attach(s)
    s.find_candidates()
    ptr, n = s.local_junctions()
    assert n == nj
    s.close()


# ---- the twopaco command line (host C++ over the C ABI) ------------------------------------------
CLI = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "twopaco_b200", "bin", "twopaco")


@pytest.mark.skipif(not os.path.exists(CLI), reason="CLI not built")
def test_cli_example_and_reference_graphdump(tmp_path, golden):
    """`twopaco -f 20 -k 11 example.fa -o example.dbg` (example/README.md:6), then the UNMODIFIED
    reference graphdump reads our file: same records as example/example.seq (canonically)."""
    import subprocess
    from tests.cases import GOLDEN_DIR
    out = tmp_path / "example.dbg"
    p = subprocess.run([CLI, "-f", "20", "-k", "11", str(GOLDEN_DIR / "example.fa"), "-o", str(out), "--tmpdir", str(tmp_path)],
                       capture_output=True, text=True)
    assert p.returncode == 0, p.stderr
    assert "Distinct junctions = 7" in p.stdout and "True marks count: 16" in p.stdout
    img = out.read_bytes()
    assert canon_md5(img) == golden["example_k11"]["canon_md5"]
    if O.have_reference():
        q = subprocess.run([str(O.REF_GRAPHDUMP), "-f", "seq", "-k", "11", str(out)], capture_output=True, text=True)
        assert q.returncode == 0, q.stderr
        rows = [tuple(int(x) for x in line.split()) for line in q.stdout.strip().splitlines()]
        seq, pos, ids = O.decode(img)
        assert rows == list(zip(seq.tolist(), pos.tolist(), ids.tolist()))
        assert [r[:2] for r in rows] == [tuple(r[:2]) for r in golden["example_k11"]["canon_stream"]]


@pytest.mark.skipif(not os.path.exists(CLI), reason="CLI not built")
def test_cli_errors_and_flags(tmp_path):
    import subprocess
    fa = tmp_path / "x.fa"
    fa.write_bytes(b">a\nACGTACGTTGCATGCATGCAAGCTTGACC\n>b\nACGTACGTTGCATGGATGCAAGCTTGACC\n")
    run = lambda *a: subprocess.run([CLI, *a], capture_output=True, text=True, cwd=tmp_path)
    assert run("-k", "5", str(fa)).returncode == 1                       # -f xor --filtermemory required
    assert run("-k", "4", "-f", "16", str(fa)).returncode == 1           # odd k
    r = run("-k", "641", "-f", "16", str(fa))
    assert r.returncode == 1 and "K is too big" in r.stderr
    r = run("-k", "5", "-f", "16", str(tmp_path / "missing.fa"))
    assert r.returncode == 1 and "Can't open file" in r.stderr
    r = run("--kvalue=5", "--filtermemory", "0.001", "-q", "3", "-r", "2", "-t", "2", "-a", "100", str(fa))
    assert r.returncode == 0 and (tmp_path / "de_bruijn.bin").exists()    # default output name
    ref, nj, _ = O.find_junctions(O.parse_fasta(str(fa)), 5, 100)
    assert (tmp_path / "de_bruijn.bin").read_bytes() == ref


@pytest.mark.skipif(not os.path.exists(CLI), reason="CLI not built")
def test_cli_selftest_mode(tmp_path):
    """`twopaco --test` (constructor.cpp:145-149): 10 random cases x k=3..9 x rounds 1..4 against a
    brute-force finder, driving the GPU path through CreateEnumerator/GetId."""
    import subprocess
    dummy = tmp_path / "dummy.fa"
    dummy.write_bytes(b">d\nACGT\n")
    r = subprocess.run([CLI, "--test", "-f", "20", "--tmpdir", str(tmp_path), str(dummy)], capture_output=True, text=True,
                       timeout=600)
    assert r.returncode == 0, r.stderr[-2000:]
    assert r.stderr.count("PASSED") == 10 and "FAILED" not in r.stderr


# ---- binned filter passes (tpc_bin.cuh): forced on small inputs via the debug environment knobs --------
@pytest.mark.parametrize("slice_log2,buffer_mb", [(13, 0), (12, 1), (16, 1), (10, 0)])
@pytest.mark.parametrize("name", ["family_k25", "family_k63", "family_k127", "family_k603", "selftest_s3_k9", "edge_mixed_k5",
                                  "family_seam_k25"])
def test_binned_filter_passes_match_golden(name, slice_log2, buffer_mb, golden, monkeypatch):
    """Partition-by-filter-slice path: single wave (records shared by fill and query), several
    waves (re-binned per pass, tiny buffer), many / few slices, overflow of skewed slices."""
    spec, g = CASES[name], golden[name]
    f = spec.get("f", 24)
    if not 1 <= f - 3 - slice_log2 <= 8:
        f = slice_log2 + 3 + 6
    monkeypatch.setenv("TPC_FILTER_MODE", "binned")
    monkeypatch.setenv("TPC_SLICE_LOG2", str(slice_log2))
    if buffer_mb:
        monkeypatch.setenv("TPC_BIN_BUFFER_MB", str(buffer_mb))
    with case_files(spec) as (paths, _, _):
        recs = api.read_fasta(paths)
    img, st = api.junctions_host(api.pack_records(recs), k=spec["k"], filter_bits=f, q=spec.get("q", 5))
    assert st.ms_bin > 0, "binned path was not taken"
    assert canon_md5(bytes(img)) == g["canon_md5"]
    assert st.junctions == g["distinct_junctions"]


@pytest.mark.parametrize("name", ["family_k29", "family_k33", "family_k65", "family_k75", "family_k97",
                                  "family_k129", "family_k257", "family_k331", "family_k603"])
def test_kmer_word_count_boundaries_on_the_sharded_binned_path(name, golden, monkeypatch):
    """k at the word-count boundaries of the packed k-mer (and at the limits of the one-window extraction of
    k_bin_list / the ownership window of k_own) through k_own + k_bin_list + the apply kernels, 3 sub-rounds."""
    spec, g = CASES[name], golden[name]
    monkeypatch.setenv("TPC_FILTER_MODE", "binned")
    monkeypatch.setenv("TPC_SLICE_LOG2", "12")
    monkeypatch.setenv("TPC_SUBROUNDS", "3")
    with case_files(spec) as (paths, _, _):
        recs = api.read_fasta(paths)
    img, st = api.junctions_host(api.pack_records(recs), k=spec["k"], filter_bits=21, q=4)
    assert st.ms_bin > 0 and st.sub_rounds == 3
    assert canon_md5(bytes(img)) == g["canon_md5"] and st.junctions == g["distinct_junctions"]


def test_binned_rounds_and_shards(monkeypatch, golden):
    monkeypatch.setenv("TPC_FILTER_MODE", "binned")
    monkeypatch.setenv("TPC_SLICE_LOG2", "12")
    spec, g = CASES["family_k25"], golden["family_k25"]
    with case_files(spec) as (paths, _, _):
        recs = api.read_fasta(paths)
    img, st = api.junctions_host(api.pack_records(recs), k=25, filter_bits=20, q=3, rounds=3)
    assert st.ms_bin > 0 and canon_md5(bytes(img)) == g["canon_md5"]
    # skew: a genome that is one k-mer repeated sends every record to a single slice: the slice's array and the overflow
    # list run over, the round is re-binned into arrays sized from the exact per-slice counts (no fall-back to the direct kernels)
    rep = [b"ACGTTGCA" * 40_000, b"ACGTTGCA" * 30_000 + b"T"]
    ref, nj, _ = O.find_junctions(rep, 25)
    img, st = api.junctions_host(api.pack_records(rep), k=25, filter_bits=22, q=5)
    assert bytes(img) == ref
    assert st.skew_rebins == 1 and st.bin_waves == 1
    # a repeat-rich family (5 % microsatellite / poly-A runs inside otherwise random genomes), with sub-rounds
    rich = synth.founder_family(4711, 5, 2, 60_000, 0.01)
    rich = [r[:20_000] + b"A" * 1500 + r[20_000:40_000] + b"AC" * 800 + r[40_000:] for r in rich]
    ref, nj, _ = O.find_junctions(rich, 25)
    monkeypatch.setenv("TPC_SUBROUNDS", "2")
    img, st = api.junctions_host(api.pack_records(rich), k=25, filter_bits=20, q=3)
    assert bytes(img) == ref and st.ms_bin > 0 and st.junctions == nj


@pytest.mark.parametrize("sub,user_rounds", [(2, 1), (3, 1), (5, 2), (9, 2)])
@pytest.mark.parametrize("name", ["family_k25", "family_k63", "family_k127", "selftest_s3_k9", "family_seam_k25"])
def test_sub_rounds_share_one_ownership_scan(name, sub, user_rounds, golden, monkeypatch):
    """Sub-rounds (automatic when one round's records do not fit HBM; forced here): the ownership of all
    the rounds of a GPU comes from ONE k_own scan as bit planes (2, 2, 4 planes; 18 rounds: one scan per
    round).  Like -r, the split must not change the result."""
    spec, g = CASES[name], golden[name]
    monkeypatch.setenv("TPC_FILTER_MODE", "binned")
    monkeypatch.setenv("TPC_SLICE_LOG2", "12")
    monkeypatch.setenv("TPC_SUBROUNDS", str(sub))
    monkeypatch.setenv("TPC_PIPELINE", "1")
    with case_files(spec) as (paths, _, _):
        recs = api.read_fasta(paths)
    img, st = api.junctions_host(api.pack_records(recs), k=spec["k"], filter_bits=20, q=4, rounds=user_rounds)
    assert st.ms_bin > 0 and st.sub_rounds == sub
    assert canon_md5(bytes(img)) == g["canon_md5"] and st.junctions == g["distinct_junctions"]
    # opt-in pipelining: rounds <= 15 per GPU are pipelined (round r+1 binned on a second stream beside the fill of round r) ...
    assert (st.ms_bin_overlapped > 0) == (sub * user_rounds <= 15)
    # ... by default they run in sequence from one scratch: same result
    monkeypatch.setenv("TPC_PIPELINE", "0")
    img2, st2 = api.junctions_host(api.pack_records(recs), k=spec["k"], filter_bits=20, q=4, rounds=user_rounds)
    assert st2.ms_bin_overlapped == 0 and bytes(img2) == bytes(img)


def test_sub_rounds_direct_path_and_shards(monkeypatch, golden):
    """Sub-rounds on the direct kernels (ownership decided inline) and combined with hash-range shards."""
    import torch
    spec, g = CASES["family_k25"], golden["family_k25"]
    with case_files(spec) as (paths, _, _):
        recs = api.read_fasta(paths)
    gen = api.pack_records(recs)
    monkeypatch.setenv("TPC_SUBROUNDS", "3")
    monkeypatch.setenv("TPC_FILTER_MODE", "direct")
    img, st = api.junctions_host(gen, k=25, filter_bits=18, q=3)
    assert st.ms_bin == 0 and canon_md5(bytes(img)) == g["canon_md5"]
    monkeypatch.setenv("TPC_FILTER_MODE", "binned")
    monkeypatch.setenv("TPC_SLICE_LOG2", "12")
    total = 0
    for i in range(2):
        s = api.Session(k=25, filter_bits=20, shard_index=i, shard_count=2)
        s.set_genome_host(gen)
        s.find_candidates()
        total += s.local_junctions()[1]
        assert s.stats().sub_rounds == 3
        s.close()
    assert total == g["distinct_junctions"]


@pytest.mark.parametrize("mode", ["direct", "binned"])
def test_host_buffer_run_with_chunked_upload(mode, monkeypatch):
    """The multi-GPU host-buffer entry point (dist.sharded_run_host) on ONE GPU: the packed genome is uploaded
    chunk by chunk on side streams and the session starts on the chunks that have arrived
    (tpc_session_add_genome_event).  The image must be byte-identical to the oracle's."""
    from twopaco_b200 import dist as tdist
    monkeypatch.setenv("TPC_FILTER_MODE", mode)
    monkeypatch.setenv("TPC_SLICE_LOG2", "12")
    recs = synth.founder_family(seed=77, genomes=5, records_per_genome=3, record_len=40_000, p=0.01, n_runs=2) + [b"ACGTACG", b""]
    g = api.pack_records(recs)
    ref, nj, nm = O.find_junctions(recs, 25)
    for n_chunks in (1, 5, 64):
        sh = tdist.host_shard(g.codes, g.n_mask, g.n_positions, g.rec_start, g.rec_len, 0, 1, n_chunks=n_chunks)
        assert 1 <= sh.plan.n_chunks <= n_chunks and (n_chunks == 1 or sh.plan.n_chunks > 1)
        info, out_host, _ = tdist.sharded_run_host(sh, 0, 1, 25, 22)
        assert info["junctions"] == nj and info["records"] == nm and info["slice_offset"] == 0
        assert out_host[:info["slice_bytes"]].numpy().tobytes() == ref


def test_sparse_n_mask_upload(monkeypatch):
    """tpc_junctions_host sets the uniform 64 KiB blocks of the n-mask on the device instead of copying them (all-zero inside the
    sequences, all-one inside long N runs) -- same image as copying everything, and the same as the oracle's; several upload chunks."""
    fam = synth.founder_family(seed=91, genomes=3, records_per_genome=2, record_len=1_500_000, p=0.002, n_runs=2)
    recs = fam[:2] + [b"ACGTTGCATGCAAGCTTGACCATGCAT" + b"N" * 1_300_000 + fam[2][:400_000]] + fam[3:] + [b"N" * 700_000]
    g = api.pack_records(recs)
    ref, nj, nm = O.find_junctions(recs, 25)
    img, st = api.junctions_host(g, k=25, filter_bits=26)
    total = g.codes.nbytes + g.n_mask.nbytes
    assert bytes(img) == ref and st.junctions == nj
    assert g.codes.nbytes < st.h2d_bytes <= total - 6 * 65536, (st.h2d_bytes, total)   # >= 3 all-zero and >= 3 all-one blocks
    monkeypatch.setenv("TPC_SPARSE_MASK", "0")
    img0, st0 = api.junctions_host(g, k=25, filter_bits=26)
    assert bytes(img0) == ref and st0.h2d_bytes == total


def test_genome_events_are_validated():
    recs = [b"ACGTTGCATGCATGCATTTGACCA" * 50]
    g = api.pack_records(recs)
    import torch
    codes = torch.from_numpy(g.codes.view(np.int64)).cuda()
    nmask = torch.from_numpy(g.n_mask.view(np.int64)).cuda()
    ev = torch.cuda.Event()
    ev.record()
    s = api.Session(k=11, filter_bits=20)
    with pytest.raises(api.TpcError):
        s.add_genome_event(0, ev.cuda_event)                  # no genome yet
    s.set_genome_device(codes.data_ptr(), nmask.data_ptr(), g.n_positions, g.rec_start, g.rec_len, keep=(codes, nmask))
    with pytest.raises(api.TpcError):
        s.add_genome_event(3, ev.cuda_event)                  # the first chunk starts at tile 0
    s.add_genome_event(0, ev.cuda_event, keep=ev)
    with pytest.raises(api.TpcError):
        s.add_genome_event(0, ev.cuda_event)                  # ascending tile order
    s.find_candidates()
    s.close()


def _run_mgpu_check(nproc: int, port: int, backend: str):
    import subprocess
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    r = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={nproc}",
                        "--master-addr", "127.0.0.1", "--master-port", str(port), os.path.join(root, "tests", "mgpu_check.py")],
                       capture_output=True, text=True, timeout=900, env={**os.environ, "TPC_MGPU_BACKEND": backend})
    assert r.returncode == 0, r.stdout[-3000:] + r.stderr[-3000:]
    assert r.stdout.count("mgpu ok") == 6 and r.stdout.count("mgpu host-buffer ok") == 6, r.stdout[-3000:]


def test_multi_gpu_torchrun():
    """N > 1: hash-range shards over NCCL (skipped on single-GPU boxes)."""
    import torch
    n = torch.cuda.device_count()
    if n < 2:
        pytest.skip("needs >= 2 GPUs")
    _run_mgpu_check(min(n, 4), 29533, "nccl")


def test_multi_rank_run_on_one_gpu_over_gloo():
    """The SAME multi-rank code (twopaco_b200.dist.sharded_run / sharded_run_host under torchrun: hash-range shards,
    all-gather of the junction words, OR-reduce of the candidate masks, position-sharded emit, chunked host upload +
    all-gather of the genome) with 3 ranks that share ONE GPU and exchange over gloo -- NCCL needs one GPU per rank,
    the exchanged bytes are the same.  Device-resident and host-buffer entry points, direct and binned filter passes,
    k = 9 / 25 / 63, byte-identical to the oracle."""
    _run_mgpu_check(3, 29534, "gloo")


def test_image_digest_device_matches_host_and_adds_over_slices():
    """tpc_image_digest_device: equal to the numpy restatement, and the digests of disjoint slices add up to the digest of
    the whole image (what lets N GPUs prove byte-identity with one GPU without gathering the image)."""
    import torch
    recs = synth.founder_family(5150, 6, 2, 60_000, 0.01, n_runs=2)
    img, _ = api.junctions_host(api.pack_records(recs), k=25, filter_bits=24)
    img = bytes(img)
    dev = torch.frombuffer(bytearray(img), dtype=torch.uint8).cuda()
    whole = api.image_digest_device(dev.data_ptr(), len(img), 0)
    assert whole == api.image_digest_host(img)
    cut = (len(img) // 36) * 12
    parts = [api.image_digest_device(dev.data_ptr(), cut, 0), api.image_digest_device(dev.data_ptr() + cut, len(img) - cut, cut)]
    assert api.add_digests(*parts) == whole
    assert api.image_digest_device(dev.data_ptr(), 0, 0) == (0, 0)
    flipped = bytearray(img)
    flipped[len(img) // 2] ^= 1
    assert api.image_digest_host(bytes(flipped)) != whole
    swapped = img[12:24] + img[:12] + img[24:]
    assert swapped == img or api.image_digest_host(swapped) != whole


def test_host_generator_matches_device_generator():
    """tools/benchutil/hostsynth.py (the sample source of `bench.py --impl reference`) == the on-device generator."""
    from tools.benchutil import hostsynth
    dg = benchutil.synth_family_device(0x4855, 3, 2, 300_000, 0.01)
    for g, c in ((0, 0), (1, 0), (2, 1)):
        want = dg.record_ascii(g * 2 + c, 120_000)
        assert hostsynth.record_prefix(0x4855, 2, 0.01, g, c, 120_000, 300_000) == want
    full = dg.record_ascii(3)
    assert hostsynth.record_prefix(0x4855, 2, 0.01, 1, 1, 10**9, 300_000) == full


def test_headline_code_path_at_100mbp_against_the_live_reference(tmp_path, monkeypatch):
    """The code path of the benchmark's headline number (C3 on one GPU: k_own with 3 ownership planes, four sub-rounds of
    k_bin_list + k_apply_fill / k_apply_query over 64 MiB slices, the slices cleared by the query kernels, merged exact pass
    from the mark list, cached ids in the emit) on 105 Mbp of the same kind of input (7 genomes, 0.1 % divergence, k = 25),
    against the UNMODIFIED reference run here with all host cores (the C oracle when the reference binary did not travel);
    the opt-in pipelined variant (five sub-rounds, round r+1 binned beside the fill of round r) and the direct kernels must
    give the same bytes."""
    dg = benchutil.synth_family_device(0x4855, 7, 1, 15_000_000, 0.001)
    recs = [dg.record_ascii(r) for r in range(7)]
    host = dg.to_host()
    monkeypatch.setenv("TPC_SUBROUNDS", "4")
    monkeypatch.setenv("TPC_FILTER_MODE", "binned")
    img, st = api.junctions_host(host, k=25, filter_bits=30, q=5)
    assert st.sub_rounds == 4 and st.ms_bin_overlapped == 0 and st.bin_waves == 1, "not the sub-round path of the benchmark"
    img = bytes(img)
    monkeypatch.setenv("TPC_SUBROUNDS", "5")
    monkeypatch.setenv("TPC_PIPELINE", "1")
    img_p, st_p = api.junctions_host(host, k=25, filter_bits=30, q=5)
    assert st_p.sub_rounds == 5 and st_p.ms_bin_overlapped > 0 and bytes(img_p) == img
    monkeypatch.delenv("TPC_PIPELINE")
    if O.have_reference():
        paths = []
        for i, r in enumerate(recs):
            p = str(tmp_path / f"g{i}.fa")
            O.write_fasta(p, [r], names=[f"g{i}_c0"])
            paths.append(p)
        ref, _ = O.run_reference(paths, 25, 30, q=5, r=1, t=os.cpu_count() or 1)
    else:
        ref, _, _ = O.find_junctions(recs, 25)
    assert len(img) == len(ref) and O.canon_equal(img, ref), "pipelined sub-round path differs from the reference"
    monkeypatch.setenv("TPC_SUBROUNDS", "0")
    monkeypatch.setenv("TPC_FILTER_MODE", "direct")
    img_d, st_d = api.junctions_host(host, k=25, filter_bits=30, q=5)
    assert st_d.bin_waves == 0 and bytes(img_d) == img
    # sharded sessions (what GPUs 0 and 3 of 4 would run: k_own + one k_bin_list pass, no sub-rounds): junction counts add up
    monkeypatch.setenv("TPC_FILTER_MODE", "binned")
    total = 0
    for i in range(4):
        s = api.Session(k=25, filter_bits=30, shard_index=i, shard_count=4)
        dg.attach(s)
        s.find_candidates()
        total += s.local_junctions()[1]
        assert s.stats().bin_waves == 1
        s.close()
    assert total == st.junctions


@pytest.mark.parametrize("name", ["family_k25", "selftest_s1_k9", "edge_mixed_k11", "family_k11_collisions"])
def test_position_identified_slots_for_short_k(name, golden, monkeypatch):
    """k <= 31 normally stores the packed k-mer inline in the table slots; the general
    position-identified slot format (used for every k > 31) must give the same result."""
    monkeypatch.setenv("TPC_INLINE_KEYS", "0")
    spec, g = CASES[name], golden[name]
    with case_files(spec) as (paths, _, _):
        recs = api.read_fasta(paths)
    img, st = api.junctions_host(api.pack_records(recs), k=spec["k"], filter_bits=spec.get("f", 24), q=spec.get("q", 5))
    assert canon_md5(bytes(img)) == g["canon_md5"] and st.junctions == g["distinct_junctions"]


@pytest.mark.parametrize("name", ["family_k25", "family_k63"])
def test_undersized_candidate_table_grows_and_redoes(name, golden, monkeypatch):
    """The candidate table is sized from a HyperLogLog estimate; exactness must not depend on it:
    a table started 64x too small overflows, is doubled and the insert pass redone until it fits."""
    monkeypatch.setenv("TPC_TABLE_SHRINK", "6")
    spec, g = CASES[name], golden[name]
    with case_files(spec) as (paths, _, _):
        recs = api.read_fasta(paths)
    img, st = api.junctions_host(api.pack_records(recs), k=spec["k"], filter_bits=14, q=2)   # many false candidates
    assert canon_md5(bytes(img)) == g["canon_md5"] and st.junctions == g["distinct_junctions"]
    assert st.candidate_kmers > 2 * st.junctions


# ---- the consumer side on the GPU: graphdump -f seq / -f group and the canonical relabelling -------------------
def _numpy_group_text(image: bytes) -> bytes:
    """graphdump.cpp:120-158 restated: classes of equal SIGNED id, members by (chr, pos), classes by first member."""
    seq, pos, ids = O.decode(image)
    classes = {}
    for i, v in enumerate(ids.tolist()):
        classes.setdefault(v, []).append(i)
    lines = []
    for members in sorted(classes.values(), key=lambda m: m[0]):
        lines.append("".join(f"{seq[i]} {pos[i]}; " for i in members) + "\n")
    return "".join(lines).encode()


@pytest.mark.parametrize("name", ["example_k11", "family_k25", "edge_mixed_k5", "edge_leading_short_k5", "family_twofiles_k25"])
def test_gpu_graphdump_seq_and_group(name, tmp_path):
    """tpc_graphdump_device / tpc_graphdump_file == the text the reference's graphdump prints (live binary when it
    travelled, else the restatement above), for our images and for a reference-style image with random ids / signs."""
    import subprocess
    spec = CASES[name]
    with case_files(spec) as (paths, _, _):
        recs = api.read_fasta(paths)
    img, _ = api.junctions_host(api.pack_records(recs), k=spec["k"], filter_bits=20)
    img = bytes(img)
    # a relabelled variant: ids permuted and signs flipped per junction, as a differently seeded reference run would write
    seq, pos, ids = O.decode(img)
    rng = np.random.default_rng(7)
    uniq = np.unique(np.abs(ids))
    perm = dict(zip(uniq.tolist(), (rng.permutation(len(uniq)) + 1000).tolist()))
    flip = {u: int(rng.integers(0, 2)) * 2 - 1 for u in uniq.tolist()}
    rec = np.frombuffer(img, dtype=O.REC_DTYPE).copy()
    sep = (rec["pos"] == O.SEP_POS) | (rec["id"] == O.SEP_ID)
    rec["id"][~sep] = [perm[abs(v)] * (1 if v > 0 else -1) * flip[abs(v)] for v in ids.tolist()]
    relabelled = rec.tobytes()
    for image in (img, relabelled):
        s2, p2, i2 = O.decode(image)
        want_seq = "".join(f"{a} {b} {c}\n" for a, b, c in zip(s2.tolist(), p2.tolist(), i2.tolist())).encode()
        assert api.graphdump(image, "seq") == want_seq
        assert api.graphdump(image, "group") == _numpy_group_text(image)
        # canonical relabelling on the GPU == oracle.canon
        canon_img, n_classes = api.canonical_image(image)
        cs, cp, cid = O.decode(canon_img)
        ws, wp, wid = O.canon(image)
        assert np.array_equal(cs, ws) and np.array_equal(cp, wp) and np.array_equal(cid, wid)
        assert n_classes == len(np.unique(np.abs(i2)))
    assert api.canonical_image(img)[0] == api.canonical_image(relabelled)[0]
    # file -> file, against the unmodified reference graphdump
    path = tmp_path / "x.dbg"
    path.write_bytes(relabelled)
    for fmt in ("seq", "group", "dot"):
        out = tmp_path / f"x.{fmt}"
        api.graphdump_file(str(path), fmt, str(out))
        if O.have_reference():
            q = subprocess.run([str(O.REF_GRAPHDUMP), "-f", fmt, "-k", str(spec["k"]), str(path)], capture_output=True)
            assert q.returncode == 0, q.stderr
            assert out.read_bytes() == q.stdout, fmt
    assert api.graphdump(b"", "seq") == b"" and api.graphdump(b"", "group") == b""


@pytest.mark.parametrize("name", ["example_k11", "family_k25", "family_twofiles_k25", "gfa_long_k25", "gfa_mixed_k11", "gfa_mixed_k5"])
def test_gpu_graphdump_gfa1_gfa2_fasta(name, tmp_path, monkeypatch):
    """tpc_graphdump_gfa_file == what the reference's graphdump prints for -f gfa1 / gfa2 / fasta (graphdump.cpp:377-582): the committed
    fixtures of the unmodified binary (tests/golden/gfa_golden.json) for our image, and for a relabelled image (random ids and
    signs, as a differently seeded reference run writes them) the live binary when it travelled and the pinned restatement."""
    import hashlib
    import json
    import subprocess
    from oracle import graphdump_gfa as G
    from tests.cases import GFA_CASES, GFA_FORMATS, GOLDEN_DIR, build_case
    spec = GFA_CASES[name]
    g = json.loads((GOLDEN_DIR / "gfa_golden.json").read_text())[name]
    monkeypatch.chdir(tmp_path)
    names = []
    for fname, content in build_case(spec):
        (tmp_path / fname).write_bytes(content)
        names.append(fname)
    img, _ = api.junctions_host(api.pack_records(api.read_fasta(names)), k=spec["k"], filter_bits=20)
    img = bytes(img)
    assert hashlib.md5(img).hexdigest() == g["image_md5"], "image differs from the oracle's: the fixtures do not apply"
    seq, pos, ids = O.decode(img)
    rng = np.random.default_rng(11)
    uniq = np.unique(np.abs(ids))
    perm = dict(zip(uniq.tolist(), (rng.permutation(len(uniq)) + 7).tolist()))
    flip = {u: int(rng.integers(0, 2)) * 2 - 1 for u in uniq.tolist()}
    rec = np.frombuffer(img, dtype=O.REC_DTYPE).copy()
    sep = (rec["pos"] == O.SEP_POS) | (rec["id"] == O.SEP_ID)
    rec["id"][~sep] = [perm[abs(v)] * (1 if v > 0 else -1) * flip[abs(v)] for v in ids.tolist()]
    (tmp_path / "image.dbg").write_bytes(img)
    (tmp_path / "relabelled.dbg").write_bytes(rec.tobytes())
    for fmt, prefix in GFA_FORMATS:
        api.graphdump_gfa_file("image.dbg", fmt, spec["k"], names, prefix, "out.txt")
        text = (tmp_path / "out.txt").read_bytes()
        want = g[fmt + ("_prefix" if prefix else "")]
        assert len(text) == want["bytes"] and hashlib.md5(text).hexdigest() == want["md5"], (fmt, prefix)
        api.graphdump_gfa_file("relabelled.dbg", fmt, spec["k"], names, prefix, "out2.txt")
        text2 = (tmp_path / "out2.txt").read_bytes()
        assert text2 == G.graphdump_text(rec.tobytes(), fmt, spec["k"], names, prefix), (fmt, prefix)
        if O.REF_GRAPHDUMP.exists():
            cmd = [str(O.REF_GRAPHDUMP), "-f", fmt, "-k", str(spec["k"])] + [a for n in names for a in ("-s", n)] + (["--prefix"] if prefix else [])
            q = subprocess.run(cmd + ["relabelled.dbg"], capture_output=True)
            assert q.returncode == 0 and q.stdout == text2, (fmt, prefix)
    # the `graphdump` command line of this build prints the same bytes
    cli = os.path.join(os.path.dirname(CLI), "graphdump")
    if os.path.exists(cli):
        q = subprocess.run([cli, "-f", "gfa2", "-k", str(spec["k"])] + [a for n in names for a in ("-s", n)] + ["--prefix", "image.dbg"], capture_output=True)
        assert q.returncode == 0 and hashlib.md5(q.stdout).hexdigest() == g["gfa2_prefix"]["md5"], q.stderr
        q = subprocess.run([cli, "-f", "seq", "-k", str(spec["k"]), "image.dbg"], capture_output=True)
        assert q.returncode == 0 and q.stdout == api.graphdump(img, "seq")
        q = subprocess.run([cli, "-f", "gfa1", "-k", str(spec["k"]), "image.dbg"], capture_output=True)
        assert q.returncode == 1 and b"seqfilename" in q.stderr


def test_gpu_graphdump_gfa_errors(tmp_path, monkeypatch):
    from tests.cases import EDGE_LEADING_SHORT
    monkeypatch.chdir(tmp_path)
    (tmp_path / "x.fa").write_bytes(EDGE_LEADING_SHORT)       # sequences shorter than k have no records: graphdump.cpp:459-462
    img, _ = api.junctions_host(api.pack_records(api.read_fasta(["x.fa"])), k=5, filter_bits=16)
    (tmp_path / "x.dbg").write_bytes(bytes(img))
    with pytest.raises(api.TpcError, match="The input is corrupted"):
        api.graphdump_gfa_file("x.dbg", "gfa1", 5, ["x.fa"], False, "o.txt")
    with pytest.raises(api.TpcError, match="format must be"):
        api.graphdump_gfa_file("x.dbg", "seq", 5, ["x.fa"], False, "o.txt")
    with pytest.raises(api.TpcError, match="Can't open file"):
        api.graphdump_gfa_file("x.dbg", "gfa2", 5, ["missing.fa"], False, "o.txt")
    (tmp_path / "big.dbg").write_bytes(np.array([(0, 1 << 31), (7, 5)], dtype=O.REC_DTYPE).tobytes())
    (tmp_path / "y.fa").write_bytes(b">y\nACGTACGTTGCATGCATGCAAGC\n")
    with pytest.raises(api.TpcError, match="A vertex id is too large"):
        api.graphdump_gfa_file("big.dbg", "gfa1", 5, ["y.fa"], False, "o.txt")
    (tmp_path / "empty.dbg").write_bytes(b"")
    api.graphdump_gfa_file("empty.dbg", "gfa1", 5, ["y.fa"], False, "o.txt")
    assert (tmp_path / "o.txt").read_bytes() == b"H\tVN:Z:1.0\nS\ty\t*\tUR:Z:y.fa\n"


DROPIN = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "oracle", "_ref", "twopaco_dropin")


@pytest.mark.skipif(not os.path.exists(DROPIN), reason="oracle/_ref/twopaco_dropin was not built (no /root/reference at build time)")
def test_reference_cli_sources_compile_and_run_against_the_library(tmp_path, golden):
    """Drop-in proof: oracle/_ref/twopaco_dropin is the reference's UNMODIFIED constructor.cpp + test.cpp (TCLAP command
    line, --test harness with its own brute-force junction finder and GetId checks) compiled with only
    vertexenumerator.{h,cpp} replaced by twopaco_b200/host's and linked with libtwopaco_b200.so (oracle/Makefile)."""
    import subprocess
    from tests.cases import GOLDEN_DIR
    out = tmp_path / "example.dbg"
    p = subprocess.run([DROPIN, "-f", "20", "-k", "11", str(GOLDEN_DIR / "example.fa"), "-o", str(out), "--tmpdir", str(tmp_path)],
                       capture_output=True, text=True)
    assert p.returncode == 0, p.stderr
    assert "Distinct junctions = 7" in p.stdout
    assert canon_md5(out.read_bytes()) == golden["example_k11"]["canon_md5"]
    dummy = tmp_path / "dummy.fa"
    dummy.write_bytes(b">d\nACGT\n")
    r = subprocess.run([DROPIN, "--test", "-f", "20", "--tmpdir", str(tmp_path), str(dummy)], capture_output=True, text=True, timeout=900)
    assert r.returncode == 0, r.stderr[-2000:]
    assert (r.stdout + r.stderr).count("PASSED") == 10 and "FAILED" not in r.stdout + r.stderr


@pytest.mark.parametrize("name", ["family_k25", "family_k63", "family_seam_k25"])
def test_mark_list_variants(name, golden, monkeypatch):
    """The exact pass reads the candidates from the mark list the binned query kernels append to (sparse marks: hash-range
    shards) or walks the mask (list off, or overflowed -> pass redone from the mask).  Same bytes either way."""
    spec, g = CASES[name], golden[name]
    with case_files(spec) as (paths, _, _):
        recs = api.read_fasta(paths)
    gen = api.pack_records(recs)
    monkeypatch.setenv("TPC_FILTER_MODE", "binned")
    monkeypatch.setenv("TPC_SLICE_LOG2", "12")
    base, st0 = api.junctions_host(gen, k=spec["k"], filter_bits=22, q=5)
    base = bytes(base)
    assert canon_md5(base) == g["canon_md5"] and st0.ms_bin > 0
    for env in ({"TPC_MARK_LIST": "0"}, {"TPC_MARK_LIST": "1"}, {"TPC_MARK_LIST": "1", "TPC_MARK_LIST_CAP": "600"},
                {"TPC_MARK_LIST": "1", "TPC_MARK_LIST_FORCE": "1", "TPC_SUBROUNDS": "3"},
                {"TPC_MARK_LIST": "1", "TPC_MARK_LIST_CAP": "600", "TPC_MARK_LIST_FORCE": "1"},
                {"TPC_MARK_LIST": "1", "TPC_MARK_LIST_CAP": "1200", "TPC_MARK_LIST_FORCE": "1", "TPC_SUBROUNDS": "2"}):
        for k_, v in env.items():
            monkeypatch.setenv(k_, v)
        img, st = api.junctions_host(gen, k=spec["k"], filter_bits=22, q=5)
        assert bytes(img) == base and st.junctions == st0.junctions, env
        for k_ in env:
            monkeypatch.delenv(k_)


# ---- the C++ multi-GPU driver (tpc_multi.cpp): one process, one host thread per GPU, NCCL between the stages ----------
def _visible_gpus():
    import torch
    return torch.cuda.device_count()


@pytest.mark.parametrize("mode", ["direct", "binned"])
def test_cxx_multi_gpu_host_buffers(mode, monkeypatch):
    """tpc_multi_junctions_host on all visible GPUs (up to 4): 1/N upload + NCCL all-gather of the genome, hash-range shards,
    all-gather of the junction words, reduce-scatter of the masks, position-sharded emit -- byte-identical to the oracle."""
    n = min(_visible_gpus(), 4)
    if n < 2:
        pytest.skip("needs >= 2 GPUs")
    monkeypatch.setenv("TPC_FILTER_MODE", mode)
    monkeypatch.setenv("TPC_SLICE_LOG2", "12")
    mg = api.MultiGpu(n)
    try:
        for k, f, recs in ((25, 24, synth.founder_family(91, 6, 3, 60_000, 0.01, n_runs=2) + [b"ACG", b""]),
                           (63, 22, synth.founder_family(92, 4, 2, 50_000, 0.02, n_runs=1)),
                           (9, 20, synth.reference_selftest_set(93)), (11, 16, [b"ACGTACGTACG"]), (11, 16, [])):
            ref, nj, nm = O.find_junctions(recs, k)
            img, st = mg.junctions_host(api.pack_records(recs), k=k, filter_bits=f)
            assert bytes(img) == ref, (k, mode)
            assert st.junctions == nj and st.occurrences == nm
            # image left on the GPUs: size and digest only
            d, nbytes, _ = mg.junctions_digest(api.pack_records(recs), k=k, filter_bits=f)
            assert nbytes == len(ref) and d == api.image_digest_host(ref)
            if k <= 31:   # position-windowed shards: 1/N of every window per GPU + all-gather; emit from private windows
                monkeypatch.setenv("TPC_WINDOW_TILES", "3")
                img, st = mg.junctions_host(api.pack_records(recs), k=k, filter_bits=f, rounds=2)
                assert bytes(img) == ref and st.junctions == nj, ("windowed", k, mode)
                d, nbytes, _ = mg.junctions_digest(api.pack_records(recs), k=k, filter_bits=f)
                assert nbytes == len(ref) and d == api.image_digest_host(ref)
                monkeypatch.delenv("TPC_WINDOW_TILES")
    finally:
        mg.close()


@pytest.mark.skipif(not os.path.exists(CLI), reason="CLI not built")
def test_cli_uses_several_gpus(tmp_path, monkeypatch):
    """`twopaco` with TPC_GPUS=N (what it does on its own for inputs >= 2^27 positions): same file as on one GPU."""
    import subprocess
    n = min(_visible_gpus(), 4)
    if n < 2:
        pytest.skip("needs >= 2 GPUs")
    recs = synth.founder_family(808, 6, 2, 120_000, 0.01, n_runs=2)
    fa = tmp_path / "in.fa"
    O.write_fasta(str(fa), recs)
    outs = []
    for gpus in (1, n):
        out = tmp_path / f"o{gpus}.bin"
        p = subprocess.run([CLI, "-k", "25", "-f", "24", str(fa), "-o", str(out), "--tmpdir", str(tmp_path)], capture_output=True, text=True,
                           env={**os.environ, "TPC_GPUS": str(gpus), "TPC_VERBOSE": "1"})
        assert p.returncode == 0, p.stderr
        assert f"{gpus} GPU(s)" in p.stderr
        outs.append(out.read_bytes())
    ref, _, _ = O.find_junctions(recs, 25)
    assert outs[0] == ref and outs[1] == ref


# ---- position-windowed sessions (tpc_windowed.inl): inputs larger than HBM stream through it window by window ----------
@pytest.mark.parametrize("window_tiles,rounds", [(1, 1), (3, 1), (7, 3), (1000, 2)])
@pytest.mark.parametrize("name", ["family_k25", "family_seam_k25", "family_k29", "selftest_s3_k9", "edge_mixed_k5", "family_twofiles_k25"])
def test_windowed_run_equals_resident_run(name, window_tiles, rounds, golden, monkeypatch):
    """TPC_WINDOW_TILES forces the driver BASELINE config 5 needs (genome, masks and planes never resident; filter, table
    and index resident; every definite k-mer position resolved against the index in the emit) on inputs small enough to
    compare: byte-identical to the resident run and to the reference's golden canonical stream, for windows of 1, 3, 7
    tiles (k-mers, records and sequence ends straddle them) and one window holding everything."""
    spec, g = CASES[name], golden[name]
    with case_files(spec) as (paths, _, _):
        recs = api.read_fasta(paths)
    gen = api.pack_records(recs)
    base, st0 = api.junctions_host(gen, k=spec["k"], filter_bits=20, q=4, rounds=rounds)
    monkeypatch.setenv("TPC_WINDOW_TILES", str(window_tiles))
    img, st = api.junctions_host(gen, k=spec["k"], filter_bits=20, q=4, rounds=rounds)
    assert bytes(img) == bytes(base), "windowed image differs from the resident one"
    assert canon_md5(bytes(img)) == g["canon_md5"]
    assert st.junctions == st0.junctions == g["distinct_junctions"] and st.occurrences == st0.occurrences


def test_host_assembled_family_equals_device_family():
    """tools/benchutil synth_family_host (C5-size inputs: generated in groups of genomes, packed at arbitrary bit offsets
    straight into host memory) == synth_family_device packed in one piece."""
    a = benchutil.synth_family_device(0x4831, 5, 3, 100_003, 0.01).to_host()
    for group in (1, 2, 5):
        b = benchutil.synth_family_host(0x4831, 5, 3, 100_003, 0.01, group=group, pin=False)
        assert b.n_positions == a.n_positions and np.array_equal(b.rec_start, a.rec_start) and np.array_equal(b.rec_len, a.rec_len)
        assert np.array_equal(b.codes, a.codes) and np.array_equal(b.n_mask, a.n_mask), group


def test_windowed_run_edge_cases(monkeypatch):
    monkeypatch.setenv("TPC_WINDOW_TILES", "2")
    for recs in ([], [b""], [b"ACG"], [b"N" * 40], [b"ACGTACGTACG"], [b"", b"ACGTTGCAAGC", b""],
                 synth.founder_family(17, 3, 2, 30_000, 0.02, n_runs=3) + [b"ACG", b""]):
        ref, nj, nm = O.find_junctions(recs, 11)
        img, st = api.junctions_host(api.pack_records(recs), k=11, filter_bits=16)
        assert bytes(img) == ref
        assert st.junctions == nj and st.occurrences == nm
    # a candidate table that starts far too small doubles and the round is redone
    monkeypatch.setenv("TPC_WINDOW_TABLE_LOG2", "4")
    recs = synth.founder_family(18, 4, 2, 40_000, 0.02)
    ref, nj, _ = O.find_junctions(recs, 25)
    img, st = api.junctions_host(api.pack_records(recs), k=25, filter_bits=18, rounds=2)
    assert bytes(img) == ref and st.junctions == nj
    monkeypatch.delenv("TPC_WINDOW_TABLE_LOG2")
    # longer k-mers keep position-identified slots, which read the genome back: not available windowed
    with pytest.raises(api.TpcError, match="k <= 31"):
        api.junctions_host(api.pack_records(recs), k=63, filter_bits=18)


@pytest.mark.parametrize("name", ["family_k25", "family_k63", "selftest_s2_k9", "edge_mixed_k5", "family_seam_k25"])
def test_list_driven_direct_passes(name, golden, monkeypatch):
    """Hash-range rounds / shards on the DIRECT filter kernels: owned positions compacted per tile and processed densely
    (k_direct_list; default) or every position rolled with an inline ownership test (k_fill / k_query, TPC_DIRECT_LIST=0)."""
    spec, g = CASES[name], golden[name]
    with case_files(spec) as (paths, _, _):
        recs = api.read_fasta(paths)
    gen = api.pack_records(recs)
    monkeypatch.setenv("TPC_FILTER_MODE", "direct")
    images = []
    for flag in ("1", "0"):
        monkeypatch.setenv("TPC_DIRECT_LIST", flag)
        img, st = api.junctions_host(gen, k=spec["k"], filter_bits=18, q=3, rounds=3)
        assert st.ms_bin == 0 and canon_md5(bytes(img)) == g["canon_md5"] and st.junctions == g["distinct_junctions"]
        images.append(bytes(img))
    assert images[0] == images[1]
    # shards: both shards' junction counts add up
    monkeypatch.setenv("TPC_DIRECT_LIST", "1")
    total = 0
    for i in range(3):
        s = api.Session(k=spec["k"], filter_bits=18, shard_index=i, shard_count=3)
        s.set_genome_host(gen)
        s.find_candidates()
        total += s.local_junctions()[1]
        s.close()
    assert total == g["distinct_junctions"]
