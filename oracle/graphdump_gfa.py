"""TEST INFRASTRUCTURE ONLY -- a sequential Python restatement of `graphdump -f gfa1 | gfa2 | fasta`
(/root/reference/src/graphdump/graphdump.cpp:47-113 Segment, :175-203 ReadInputSequences, :205-374 the generators,
:377-466 GenerateGfaOutput, :468-582 GenerateFastaOutput; sequences as ChrReader / StreamFastaParser read them,
src/common/streamfastaparser.h:140-182, streamfastaparser.cpp:29-94).

Pinned: tests/test_oracle.py compares it with tests/golden/gfa_golden.json, i.e. with the output of the UNMODIFIED
reference binary on the same images and FASTA files (tests/golden/make_gfa_golden.py).  The GPU tests use it as the checker
for images the fixtures do not cover (relabelled ids).  Never imported by the product."""
from __future__ import annotations

import numpy as np

SEP_POS = 0xFFFFFFFF
SEP_ID = np.iinfo(np.int64).max
VALID = set(b"ACGTURYKMSWBDHWNXV")            # dnachar.cpp:9-11
RESERVED_PATH0 = 1 << 34                       # graphdump.cpp:42 (ID_POWER 35)
MAX_JUNCTION_ID = 1 << 31                      # :44
_REV = {ord("A"): "T", ord("T"): "A", ord("C"): "G", ord("G"): "C"}


class GraphdumpError(RuntimeError):
    pass


def read_sequences(paths: list[str]):
    """-> (upper-cased sequences, header tokens, file of each record); a header line without a token keeps the previous one."""
    seqs, headers, files = [], [], []
    current = ""
    for path in paths:
        data = open(path, "rb").read()
        i, n = 0, len(data)
        while i < n:
            if data[i] != ord(">"):
                raise GraphdumpError(f"The FASTA header should start with a '>', started with '{chr(data[i])}'")
            j = data.find(b"\n", i + 1)
            if j >= 0:
                tok = data[i + 1:j].split()
                if tok:
                    current = tok[0].decode("latin-1")
                i = j + 1
            else:
                i = n
            j = data.find(b">", i)
            j = n if j < 0 else j
            body = bytes(c for c in data[i:j].upper() if not chr(c).isspace())
            bad = [c for c in body if c not in VALID]
            if bad:
                raise GraphdumpError(f"Found an invalid character '{chr(bad[0])}' in sequence {current}")
            seqs.append(body); headers.append(current); files.append(path)
            i = j
    return seqs, headers, files


def _records(image: bytes):
    rec = np.frombuffer(image[:len(image) // 12 * 12], dtype=np.dtype([("pos", "<u4"), ("id", "<i8")]))
    chrom = 0
    for pos, jid in zip(rec["pos"].tolist(), rec["id"].tolist()):
        if pos == SEP_POS or jid == SEP_ID:     # junctionapi.h:94
            chrom += 1
        else:
            yield chrom, pos, jid


def _sign(v: int) -> str:
    return "+" if v >= 0 else "-"


def _gfa2_pos(pos: int, length: int) -> str:
    return f"{pos}$" if pos == length else f"{pos}"


def _revcomp(s: bytes) -> str:
    return "".join(_REV.get(c, "N") for c in reversed(s))


def graphdump_text(image: bytes, fmt: str, k: int, seq_paths: list[str], prefix: bool = False) -> bytes:
    seqs, headers, files = read_sequences(seq_paths)
    names = [("s0_" + h) if (prefix and fmt != "fasta") else h for h in headers]     # :183-196: the counter never advances
    file_of = {}
    for nm, f in zip(names, files):
        file_of[nm] = f
    out = []
    if fmt == "gfa1":
        out.append("H\tVN:Z:1.0\n")
        out += [f"S\t{nm}\t*\tUR:Z:{file_of[nm]}\n" for nm in names]
    elif fmt == "gfa2":
        out.append("H\tVN:Z:2.0\n")
    reserved = RESERVED_PATH0
    seen = set()
    path: list[int] = []
    seq_id = 0
    prev_seg, prev_size = 0, -1
    begin = None

    def flush():
        if path and fmt != "fasta":
            items = [f"{abs(s)}{_sign(s)}" for s in path]
            out.append(f"P\t{names[seq_id]}\t" + ",".join(items) + "\t*\n" if fmt == "gfa1" else f"O\t{names[seq_id]}p\t" + " ".join(items) + "\n")
        path.clear()

    for end in _records(image):
        if begin is None:
            begin = end
            if begin[0] != 0 or not seqs:
                raise GraphdumpError("The input is corrupted")      # (undefined behaviour in the reference)
            continue
        if begin[0] == end[0]:
            chrom = seqs[seq_id]
            if end[1] + k > len(chrom) or end[1] <= begin[1]:
                raise GraphdumpError("The input is corrupted")
            pos_edge, neg_edge = chr(chrom[begin[1] + k]), _REV.get(chrom[end[1] - 1], "N")
            ab, ae = abs(begin[2]), abs(end[2])
            if ab >= MAX_JUNCTION_ID or ae >= MAX_JUNCTION_ID:
                raise GraphdumpError("A vertex id is too large, cannot generate GFA")
            fwd = ab < ae or (ab == ae and ab > 0)
            edge, b_id = (pos_edge, begin[2]) if fwd else (neg_edge, -end[2])
            if edge == "N":
                seg = reserved
                reserved += 1
            else:
                seg = "ACGT".find(edge)                              # MakeUpChar: -1 (as size_t) for anything else
                seg |= (4 | (abs(b_id) << 3)) if b_id < 0 else (b_id << 3)
                if begin[2] != b_id:
                    seg = -seg
            path.append(seg)
            size = end[1] + k - begin[1]
            if abs(seg) not in seen:
                seen.add(abs(seg))
                raw = chrom[begin[1]:end[1] + k]
                body = raw.decode("latin-1") if seg > 0 else _revcomp(raw)
                if fmt == "gfa1":
                    out.append(f"S\t{abs(seg)}\t{body}\n")
                elif fmt == "gfa2":
                    out.append(f"S\t{abs(seg)}\t{size}\t{body}\n")
                else:
                    out.append(f">{abs(seg)}\n" + "".join(body[i:i + 80] + "\n" for i in range(0, len(body), 80)))
            if fmt == "gfa1":
                out.append(f"C\t{abs(seg)}\t{_sign(seg)}\t{names[seq_id]}\t+\t{end[1]}\n")
                if prev_seg:
                    out.append(f"L\t{abs(prev_seg)}\t{_sign(prev_seg)}\t{abs(seg)}\t{_sign(seg)}\t{k}M\n")
            elif fmt == "gfa2":
                clen = len(chrom)
                out.append(f"F\t{abs(seg)}\t{names[seq_id]}{_sign(seg)}\t0\t{size}$\t{_gfa2_pos(begin[1], clen)}\t{_gfa2_pos(end[1] + k, clen)}\t{k}M\n")
                if prev_seg:
                    p0, p1 = (prev_size - k, prev_size) if prev_seg > 0 else (0, k)
                    s0, s1 = (0, k) if seg > 0 else (size - k, size)
                    out.append(f"E\t{abs(prev_seg)}{_sign(prev_seg)}\t{abs(seg)}{_sign(seg)}\t{_gfa2_pos(p0, prev_size)}\t{_gfa2_pos(p1, prev_size)}\t"
                               f"{_gfa2_pos(s0, size)}\t{_gfa2_pos(s1, size)}\t{k}M\n")
            prev_seg, prev_size = seg, size
            begin = end
        else:
            flush()
            prev_seg = 0
            begin = end
            seq_id += 1
            if begin[0] != seq_id or seq_id >= len(seqs):
                raise GraphdumpError("The input is corrupted")       # :459-462
    if begin is not None:
        flush()
    return "".join(out).encode("latin-1")
