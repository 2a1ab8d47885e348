"""ctypes binding of libtwopaco_b200.so (include/twopaco_b200.h) -- plumbing only.

Every function here calls straight into the CUDA library; there is no Python or CPU
implementation of the path behind it.  If the library is missing, or no GPU is present,
calls fail loudly.
"""
from __future__ import annotations

import ctypes as C
import os
from pathlib import Path

import numpy as np

PKG = Path(__file__).resolve().parent
LIB_PATH = PKG / "libtwopaco_b200.so"

ABUNDANCE_MAX = 2**64 - 1
INVALID_VERTEX = 2**63 - 1
TILE_POSITIONS = 8192


class TpcError(RuntimeError):
    pass


class Params(C.Structure):
    _fields_ = [("k", C.c_uint32), ("filter_bits", C.c_uint32), ("q", C.c_uint32), ("rounds", C.c_uint32),
                ("abundance", C.c_uint64), ("shard_index", C.c_uint32), ("shard_count", C.c_uint32),
                ("seed", C.c_uint64)]


class Stats(C.Structure):
    _fields_ = [("positions", C.c_uint64), ("candidate_marks", C.c_uint64), ("candidate_kmers", C.c_uint64),
                ("junctions", C.c_uint64), ("occurrences", C.c_uint64), ("stubs", C.c_uint64),
                ("out_bytes", C.c_uint64), ("filter_edges_set", C.c_uint64),
                ("ms_bin", C.c_float), ("ms_fill", C.c_float), ("ms_query", C.c_float), ("ms_insert", C.c_float), ("ms_classify", C.c_float),
                ("ms_index", C.c_float), ("ms_emit", C.c_float), ("ms_total", C.c_float),
                ("kernel_launches", C.c_uint32), ("bin_waves", C.c_uint32), ("sub_rounds", C.c_uint32),
                ("ms_bin_overlapped", C.c_float), ("ms_wall_candidates", C.c_float), ("ms_wall_index", C.c_float),
                ("ms_wall_emit", C.c_float), ("skew_rebins", C.c_uint32), ("h2d_bytes", C.c_uint64)]

    def asdict(self) -> dict:
        return {n: getattr(self, n) for n, _ in self._fields_}


class Genome(C.Structure):
    _fields_ = [("codes", C.c_void_p), ("n_mask", C.c_void_p), ("n_positions", C.c_uint64),
                ("rec_start", C.c_void_p), ("rec_len", C.c_void_p), ("n_records", C.c_uint64)]


LOG_FN = C.CFUNCTYPE(None, C.c_void_p, C.c_char_p)

# name -> (restype, argtypes); also the list test_abi checks against the header
SIGNATURES = {
    "tpc_code_words": (C.c_uint64, [C.c_uint64]),
    "tpc_mask_words": (C.c_uint64, [C.c_uint64]),
    "tpc_positions_for": (C.c_uint64, [C.c_void_p, C.c_uint64]),
    "tpc_pack_records": (C.c_int, [C.POINTER(C.c_char_p), C.c_void_p, C.c_uint64, C.c_uint32, C.c_void_p, C.c_void_p, C.c_void_p]),
    "tpc_read_fasta": (C.c_int, [C.c_char_p, C.POINTER(C.POINTER(C.c_char_p)), C.POINTER(C.POINTER(C.c_uint64)), C.POINTER(C.c_uint64)]),
    "tpc_free_records": (None, [C.POINTER(C.c_char_p), C.POINTER(C.c_uint64), C.c_uint64]),
    "tpc_ingest_fasta": (C.c_int, [C.POINTER(C.c_char_p), C.c_size_t, C.c_uint32, C.POINTER(C.c_void_p), C.POINTER(C.c_uint64),
                                   C.POINTER(C.c_void_p), C.POINTER(C.c_void_p), C.POINTER(C.c_uint64)]),
    "tpc_host_free": (None, [C.c_void_p]),
    "tpc_build": (C.c_int, [C.POINTER(C.c_char_p), C.c_size_t, C.c_uint32, C.c_uint32, C.c_uint32, C.c_uint32, C.c_uint32,
                            C.c_uint64, C.c_char_p, C.c_char_p, LOG_FN, C.c_void_p, C.POINTER(C.c_void_p)]),
    "tpc_vertices": (C.c_uint64, [C.c_void_p]),
    "tpc_get_id": (C.c_int64, [C.c_void_p, C.c_char_p]),
    "tpc_handle_stats": (C.c_int, [C.c_void_p, C.POINTER(Stats)]),
    "tpc_free": (None, [C.c_void_p]),
    "tpc_junctions_host": (C.c_int, [C.POINTER(Params), C.POINTER(Genome), C.c_void_p, C.c_uint64, C.POINTER(C.c_uint64), C.POINTER(Stats)]),
    "tpc_visible_gpus": (C.c_uint32, []),
    "tpc_multi_create": (C.c_int, [C.c_uint32, C.c_void_p, C.POINTER(C.c_void_p)]),
    "tpc_multi_destroy": (None, [C.c_void_p]),
    "tpc_multi_gpus": (C.c_uint32, [C.c_void_p]),
    "tpc_multi_junctions_host": (C.c_int, [C.c_void_p, C.POINTER(Params), C.POINTER(Genome), C.c_void_p, C.c_uint64,
                                           C.POINTER(C.c_uint64), C.POINTER(Stats)]),
    "tpc_multi_junctions_digest": (C.c_int, [C.c_void_p, C.POINTER(Params), C.POINTER(Genome), C.POINTER(C.c_uint64 * 2),
                                             C.POINTER(C.c_uint64), C.POINTER(Stats)]),
    "tpc_session_create": (C.c_int, [C.POINTER(Params), C.c_void_p, C.POINTER(C.c_void_p)]),
    "tpc_session_destroy": (None, [C.c_void_p]),
    "tpc_session_set_genome_host": (C.c_int, [C.c_void_p, C.POINTER(Genome)]),
    "tpc_session_set_genome_device": (C.c_int, [C.c_void_p, C.POINTER(Genome)]),
    "tpc_session_add_genome_event": (C.c_int, [C.c_void_p, C.c_uint64, C.c_void_p]),
    "tpc_session_find_candidates": (C.c_int, [C.c_void_p]),
    "tpc_session_local_junctions": (C.c_int, [C.c_void_p, C.POINTER(C.c_void_p), C.POINTER(C.c_uint64)]),
    "tpc_session_set_junctions": (C.c_int, [C.c_void_p, C.c_void_p, C.c_uint64]),
    "tpc_session_candidate_mask": (C.c_int, [C.c_void_p, C.POINTER(C.c_void_p), C.POINTER(C.c_uint64)]),
    "tpc_session_emit_count": (C.c_int, [C.c_void_p, C.c_uint64, C.c_uint64, C.POINTER(C.c_uint64), C.POINTER(C.c_uint64)]),
    "tpc_session_emit_write": (C.c_int, [C.c_void_p, C.c_uint64, C.c_uint64, C.c_void_p, C.c_uint64, C.POINTER(C.c_uint64), C.POINTER(C.c_uint64)]),
    "tpc_session_get_id": (C.c_int, [C.c_void_p, C.c_char_p, C.POINTER(C.c_int64)]),
    "tpc_session_stats": (C.c_int, [C.c_void_p, C.POINTER(Stats)]),
    "tpc_pack_ascii_device": (C.c_int, [C.c_void_p, C.c_uint64, C.c_void_p, C.c_void_p, C.c_void_p]),
    "tpc_image_digest_device": (C.c_int, [C.c_void_p, C.c_uint64, C.c_uint64, C.c_void_p, C.POINTER(C.c_uint64 * 2)]),
    "tpc_graphdump_device": (C.c_int, [C.c_void_p, C.c_uint64, C.c_uint32, C.c_void_p, C.POINTER(C.c_void_p), C.POINTER(C.c_uint64)]),
    "tpc_graphdump_file": (C.c_int, [C.c_char_p, C.c_char_p, C.c_char_p]),
    "tpc_graphdump_gfa_device": (C.c_int, [C.c_void_p, C.c_uint64, C.c_uint32, C.c_uint32, C.c_void_p, C.c_void_p, C.POINTER(C.c_char_p),
                                           C.c_uint64, C.c_void_p, C.POINTER(C.c_void_p), C.POINTER(C.c_uint64)]),
    "tpc_graphdump_gfa_file": (C.c_int, [C.c_char_p, C.c_char_p, C.c_uint32, C.POINTER(C.c_char_p), C.c_size_t, C.c_int, C.c_char_p]),
    "tpc_canonical_image_device": (C.c_int, [C.c_void_p, C.c_uint64, C.c_void_p, C.c_void_p, C.POINTER(C.c_uint64)]),
    "tpc_release_cached_memory": (C.c_int, []),
    "tpc_device_alloc": (C.c_int, [C.c_uint64, C.POINTER(C.c_void_p)]),
    "tpc_device_free": (None, [C.c_void_p]),
    "tpc_copy_to_host": (C.c_int, [C.c_void_p, C.c_void_p, C.c_uint64]),
    "tpc_copy_to_device": (C.c_int, [C.c_void_p, C.c_void_p, C.c_uint64]),
    "tpc_last_error": (C.c_char_p, []),
    "tpc_abi_version": (C.c_uint32, []),
}

_lib = None


def lib() -> C.CDLL:
    """Load the CUDA library; raises if it has not been built (no fallback)."""
    global _lib
    if _lib is None:
        if not LIB_PATH.exists():
            raise TpcError(f"{LIB_PATH} is missing: run `python -c 'import __graft_entry__ as g; g.build()'` "
                           "(twopaco_b200 has no CPU fallback)")
        L = C.CDLL(str(LIB_PATH))
        for name, (res, args) in SIGNATURES.items():
            fn = getattr(L, name)
            fn.restype, fn.argtypes = res, args
        _lib = L
    return _lib


def _check(rc: int) -> None:
    if rc != 0:
        raise TpcError(lib().tpc_last_error().decode(errors="replace"))


# ---------------------------------------------------------------------------------------------
# host-side packing
# ---------------------------------------------------------------------------------------------
class PackedGenome:
    """2-bit packed genome in host memory (numpy arrays; layout: include/twopaco_b200.h)."""

    def __init__(self, codes, n_mask, n_positions, rec_start, rec_len):
        self.codes, self.n_mask, self.n_positions = codes, n_mask, int(n_positions)
        self.rec_start, self.rec_len = rec_start, rec_len

    @property
    def n_records(self) -> int:
        return len(self.rec_len)

    @property
    def total_bp(self) -> int:
        return int(self.rec_len.sum())

    def struct(self) -> Genome:
        return Genome(self.codes.ctypes.data, self.n_mask.ctypes.data, self.n_positions,
                      self.rec_start.ctypes.data, self.rec_len.ctypes.data, self.n_records)


def pack_records(records: list[bytes], threads: int = 0) -> PackedGenome:
    L = lib()
    n = len(records)
    rec_len = np.array([len(r) for r in records], dtype=np.uint64)
    npos = L.tpc_positions_for(rec_len.ctypes.data, n)
    codes = np.empty(L.tpc_code_words(npos), dtype=np.uint64)
    n_mask = np.empty(L.tpc_mask_words(npos), dtype=np.uint64)
    rec_start = np.empty(max(n, 1), dtype=np.uint64)[:n]
    arr = (C.c_char_p * max(n, 1))(*records)
    _check(L.tpc_pack_records(arr, rec_len.ctypes.data, n, threads or os.cpu_count() or 1,
                              codes.ctypes.data, n_mask.ctypes.data, rec_start.ctypes.data if n else None))
    return PackedGenome(codes, n_mask, npos, rec_start, rec_len)


def read_fasta(paths: list[str]) -> list[bytes]:
    L = lib()
    recs = C.POINTER(C.c_char_p)()
    lens = C.POINTER(C.c_uint64)()
    n = C.c_uint64(0)
    try:
        for p in paths:
            _check(L.tpc_read_fasta(os.fsencode(p), C.byref(recs), C.byref(lens), C.byref(n)))
        return [C.string_at(recs[i], lens[i]) for i in range(n.value)]
    finally:
        L.tpc_free_records(recs, lens, n.value)


def ingest_fasta(paths: list[str], threads: int = 0):
    """Multi-threaded parser of tpc_build -> (ascii position layout as bytes, n_positions, rec_start, rec_len)."""
    L = lib()
    arr = (C.c_char_p * max(len(paths), 1))(*[os.fsencode(p) for p in paths])
    a, s, l = C.c_void_p(), C.c_void_p(), C.c_void_p()
    npos, nrec = C.c_uint64(), C.c_uint64()
    _check(L.tpc_ingest_fasta(arr, len(paths), threads or os.cpu_count() or 1, C.byref(a), C.byref(npos), C.byref(s), C.byref(l),
                              C.byref(nrec)))
    try:
        n = nrec.value
        layout = C.string_at(a, npos.value)
        rec_start = np.frombuffer(C.string_at(s, n * 8), dtype=np.uint64).copy()
        rec_len = np.frombuffer(C.string_at(l, n * 8), dtype=np.uint64).copy()
    finally:
        for p in (a, s, l):
            L.tpc_host_free(p)
    return layout, npos.value, rec_start, rec_len


# ---------------------------------------------------------------------------------------------
# level 2: packed host genome -> image
# ---------------------------------------------------------------------------------------------
def junctions_host(genome: PackedGenome, k: int, filter_bits: int, q: int = 5, rounds: int = 1,
                   abundance: int = ABUNDANCE_MAX, seed: int = 0, out: np.ndarray | None = None):
    """-> (image bytes as numpy uint8 view, Stats).  `out` may be a preallocated (pinned) buffer."""
    L = lib()
    prm = Params(k, filter_bits, q, rounds, abundance, 0, 1, seed)
    g = genome.struct()
    st = Stats()
    nbytes = C.c_uint64(0)
    if out is None:
        out = np.empty(max(12 * (genome.n_records * 2 + 1024), 1 << 16), dtype=np.uint8)
    rc = L.tpc_junctions_host(C.byref(prm), C.byref(g), out.ctypes.data, out.nbytes, C.byref(nbytes), C.byref(st))
    if rc == 2:  # buffer too small: retry with the exact size
        out = np.empty(nbytes.value, dtype=np.uint8)
        rc = L.tpc_junctions_host(C.byref(prm), C.byref(g), out.ctypes.data, out.nbytes, C.byref(nbytes), C.byref(st))
    _check(rc)
    return out[:nbytes.value], st


class MultiGpu:
    """N GPUs of this process (tpc_multi_*): one host thread + one hash-range shard per GPU, NCCL in between."""

    def __init__(self, n_gpus: int, devices: list[int] | None = None):
        self._m = C.c_void_p()
        arr = (C.c_int * n_gpus)(*devices) if devices else None
        _check(lib().tpc_multi_create(n_gpus, arr, C.byref(self._m)))
        self.n_gpus = n_gpus

    def junctions_host(self, genome: PackedGenome, k: int, filter_bits: int, q: int = 5, rounds: int = 1,
                       abundance: int = ABUNDANCE_MAX, seed: int = 0, out: np.ndarray | None = None):
        """-> (image bytes as numpy uint8 view, Stats summed / maxed over the shards)."""
        L = lib()
        prm = Params(k, filter_bits, q, rounds, abundance, 0, self.n_gpus, seed)
        g = genome.struct()
        st = Stats()
        nbytes = C.c_uint64(0)
        if out is None:
            out = np.empty(max(12 * (genome.n_records * 2 + 1024), 1 << 16), dtype=np.uint8)
        rc = L.tpc_multi_junctions_host(self._m, C.byref(prm), C.byref(g), out.ctypes.data, out.nbytes, C.byref(nbytes), C.byref(st))
        if rc == 2:   # buffer too small: the error message carries a lower bound; grow geometrically
            for _ in range(8):
                out = np.empty(max(out.nbytes * 4, 1 << 20), dtype=np.uint8)
                rc = L.tpc_multi_junctions_host(self._m, C.byref(prm), C.byref(g), out.ctypes.data, out.nbytes, C.byref(nbytes), C.byref(st))
                if rc != 2:
                    break
        _check(rc)
        return out[:nbytes.value], st

    def junctions_digest(self, genome: PackedGenome, k: int, filter_bits: int, q: int = 5, rounds: int = 1, seed: int = 0):
        """The same run with the image left on the GPUs -> ((digest a, digest b), image bytes, Stats)."""
        prm = Params(k, filter_bits, q, rounds, ABUNDANCE_MAX, 0, self.n_gpus, seed)
        g = genome.struct()
        st, d, nbytes = Stats(), (C.c_uint64 * 2)(), C.c_uint64(0)
        _check(lib().tpc_multi_junctions_digest(self._m, C.byref(prm), C.byref(g), C.byref(d), C.byref(nbytes), C.byref(st)))
        return (int(d[0]), int(d[1])), nbytes.value, st

    def close(self) -> None:
        if self._m:
            lib().tpc_multi_destroy(self._m)
            self._m = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


# ---------------------------------------------------------------------------------------------
# level 1: files -> file (the reference's CreateEnumerator)
# ---------------------------------------------------------------------------------------------
class VertexEnumerator:
    """Mirror of TwoPaCo::VertexEnumerator (vertexenumerator.h:23-35)."""

    def __init__(self, handle, log_text: str):
        self._h = handle
        self.log = log_text

    def GetVerticesCount(self) -> int:
        return lib().tpc_vertices(self._h)

    def GetId(self, vertex: str) -> int:
        return lib().tpc_get_id(self._h, vertex.encode())

    def stats(self) -> Stats:
        st = Stats()
        _check(lib().tpc_handle_stats(self._h, C.byref(st)))
        return st

    def close(self) -> None:
        if self._h:
            lib().tpc_free(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


def CreateEnumerator(fileName: list[str], vertexLength: int, filterSize: int, hashFunctions: int = 5, rounds: int = 1,
                     threads: int = 1, abundance: int = ABUNDANCE_MAX, tmpDirName: str = ".",
                     outFileName: str = "de_bruijn.bin") -> VertexEnumerator:
    """TwoPaCo::CreateEnumerator (vertexenumerator.h:37-46) on the current CUDA device."""
    L = lib()
    chunks: list[str] = []
    cb = LOG_FN(lambda ctx, text: chunks.append(text.decode(errors="replace")))
    arr = (C.c_char_p * len(fileName))(*[os.fsencode(p) for p in fileName])
    h = C.c_void_p()
    _check(L.tpc_build(arr, len(fileName), vertexLength, filterSize, hashFunctions, rounds, threads, abundance,
                       os.fsencode(tmpDirName), os.fsencode(outFileName), cb, None, C.byref(h)))
    return VertexEnumerator(h, "".join(chunks))


# ---------------------------------------------------------------------------------------------
# level 3: sessions
# ---------------------------------------------------------------------------------------------
class Session:
    def __init__(self, k: int, filter_bits: int, q: int = 5, rounds: int = 1, abundance: int = ABUNDANCE_MAX,
                 shard_index: int = 0, shard_count: int = 1, seed: int = 0, stream: int = 0):
        self._s = C.c_void_p()
        prm = Params(k, filter_bits, q, rounds, abundance, shard_index, shard_count, seed)
        _check(lib().tpc_session_create(C.byref(prm), C.c_void_p(stream), C.byref(self._s)))
        self._keep = None

    def set_genome_host(self, genome: PackedGenome) -> None:
        self._keep = genome
        g = genome.struct()
        _check(lib().tpc_session_set_genome_host(self._s, C.byref(g)))

    def set_genome_device(self, codes_ptr: int, n_mask_ptr: int, n_positions: int, rec_start: np.ndarray,
                          rec_len: np.ndarray, keep=None) -> None:
        self._keep = (keep, rec_start, rec_len)
        g = Genome(codes_ptr, n_mask_ptr, n_positions, rec_start.ctypes.data, rec_len.ctypes.data, len(rec_len))
        _check(lib().tpc_session_set_genome_device(self._s, C.byref(g)))

    def add_genome_event(self, tile_begin: int, cuda_event: int, keep=None) -> None:
        """Tiles from tile_begin on (up to the next event's tile) are complete once the CUDA event has completed."""
        self._events = getattr(self, "_events", []) + [keep]
        _check(lib().tpc_session_add_genome_event(self._s, tile_begin, C.c_void_p(cuda_event)))

    def find_candidates(self) -> None:
        _check(lib().tpc_session_find_candidates(self._s))

    def local_junctions(self) -> tuple[int, int]:
        p, n = C.c_void_p(), C.c_uint64()
        _check(lib().tpc_session_local_junctions(self._s, C.byref(p), C.byref(n)))
        return p.value or 0, n.value

    def set_junctions(self, dev_ptr: int, count: int) -> None:
        _check(lib().tpc_session_set_junctions(self._s, C.c_void_p(dev_ptr), count))

    def candidate_mask(self) -> tuple[int, int]:
        p, n = C.c_void_p(), C.c_uint64()
        _check(lib().tpc_session_candidate_mask(self._s, C.byref(p), C.byref(n)))
        return p.value or 0, n.value

    def emit_count(self, pos_begin: int, pos_end: int) -> tuple[int, int]:
        a, b = C.c_uint64(), C.c_uint64()
        _check(lib().tpc_session_emit_count(self._s, pos_begin, pos_end, C.byref(a), C.byref(b)))
        return a.value, b.value

    def emit_write(self, records_before: int, stubs_before: int, dev_out: int, capacity: int) -> tuple[int, int]:
        off, nb = C.c_uint64(), C.c_uint64()
        _check(lib().tpc_session_emit_write(self._s, records_before, stubs_before, C.c_void_p(dev_out), capacity,
                                            C.byref(off), C.byref(nb)))
        return off.value, nb.value

    def get_id(self, kmer: str) -> int:
        v = C.c_int64()
        _check(lib().tpc_session_get_id(self._s, kmer.encode(), C.byref(v)))
        return v.value

    def stats(self) -> Stats:
        st = Stats()
        _check(lib().tpc_session_stats(self._s, C.byref(st)))
        return st

    def close(self) -> None:
        if self._s:
            lib().tpc_session_destroy(self._s)
            self._s = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


def release_cached_memory() -> None:
    _check(lib().tpc_release_cached_memory())


def image_digest_device(dev_ptr: int, nbytes: int, image_offset: int = 0, stream: int = 0) -> tuple[int, int]:
    """Position-keyed digest of an image slice in device memory (slices' digests add up mod 2^64)."""
    d = (C.c_uint64 * 2)()
    _check(lib().tpc_image_digest_device(C.c_void_p(dev_ptr), nbytes, image_offset, C.c_void_p(stream), C.byref(d)))
    return int(d[0]), int(d[1])


def graphdump(image, fmt: str = "seq") -> bytes:
    """GPU `graphdump -f seq|group` of an image given as host bytes -> the text graphdump prints."""
    data = np.frombuffer(bytes(image), dtype=np.uint8)
    buf = DeviceBuffer(max(len(data), 16))
    if len(data):
        buf.from_host(data)
    t, n = C.c_void_p(), C.c_uint64()
    _check(lib().tpc_graphdump_device(C.c_void_p(buf.ptr), len(data), {"seq": 0, "group": 1, "dot": 2}[fmt], None, C.byref(t), C.byref(n)))
    text = DeviceBuffer.adopt(t.value, n.value)
    try:
        return text.to_host(n.value).tobytes() if n.value else b""
    finally:
        text.close()
        buf.close()


def graphdump_file(image_path: str, fmt: str, out_path: str | None = None) -> None:
    _check(lib().tpc_graphdump_file(os.fsencode(image_path), fmt.encode(), os.fsencode(out_path) if out_path else None))


def graphdump_gfa_file(image_path: str, fmt: str, k: int, seq_paths: list[str], prefix: bool = False, out_path: str | None = None) -> None:
    """GPU `graphdump -f gfa1|gfa2|fasta -k k -s <fasta>... [--prefix] image` -> text file (None = stdout)."""
    arr = (C.c_char_p * max(len(seq_paths), 1))(*[os.fsencode(p) for p in seq_paths])
    _check(lib().tpc_graphdump_gfa_file(os.fsencode(image_path), fmt.encode(), k, arr, len(seq_paths), int(prefix),
                                        os.fsencode(out_path) if out_path else None))


def canonical_image(image) -> tuple[bytes, int]:
    """Canonical relabelling of an image on the GPU -> (canonical image bytes, number of distinct |id|)."""
    data = np.frombuffer(bytes(image), dtype=np.uint8)
    a, b = DeviceBuffer(max(len(data), 16)), DeviceBuffer(max(len(data), 16))
    if len(data):
        a.from_host(data)
    n = C.c_uint64()
    try:
        _check(lib().tpc_canonical_image_device(C.c_void_p(a.ptr), len(data), None, C.c_void_p(b.ptr), C.byref(n)))
        return (b.to_host(len(data)).tobytes() if len(data) else b""), n.value
    finally:
        a.close()
        b.close()


def _fmix64(x: np.ndarray) -> np.ndarray:
    x = x ^ (x >> np.uint64(33)); x = x * np.uint64(0xff51afd7ed558ccd)
    x = x ^ (x >> np.uint64(33)); x = x * np.uint64(0xc4ceb9fe1a85ec53)
    return x ^ (x >> np.uint64(33))


def image_digest_host(image, image_offset: int = 0) -> tuple[int, int]:
    """The same digest computed with numpy from image bytes in host memory."""
    w = np.frombuffer(bytes(image), dtype="<u4").astype(np.uint64)
    a = b = 0
    with np.errstate(over="ignore"):
        for lo in range(0, len(w), 1 << 24):
            part = w[lo:lo + (1 << 24)]
            gi = np.arange(image_offset // 4 + lo + 1, image_offset // 4 + lo + 1 + len(part), dtype=np.uint64)
            a = (a + int(_fmix64((gi * np.uint64(0x9E3779B97F4A7C15)) ^ part).sum(dtype=np.uint64))) & (2**64 - 1)
            b = (b + int(_fmix64(gi * np.uint64(0xC2B2AE3D27D4EB4F) + part * np.uint64(0x165667B19E3779F9)).sum(dtype=np.uint64))) & (2**64 - 1)
    return a, b


def add_digests(*ds: tuple[int, int]) -> tuple[int, int]:
    return (sum(d[0] for d in ds) & (2**64 - 1), sum(d[1] for d in ds) & (2**64 - 1))


# ---------------------------------------------------------------------------------------------
# device-resident genomes (benchmark inputs; K0 pack)
# ---------------------------------------------------------------------------------------------
class DeviceBuffer:
    """cudaMalloc'ed buffer owned by Python (freed on close / GC)."""

    def __init__(self, nbytes: int):
        p = C.c_void_p()
        _check(lib().tpc_device_alloc(nbytes, C.byref(p)))
        self.ptr, self.nbytes = p.value, nbytes

    @classmethod
    def adopt(cls, ptr: int, nbytes: int) -> "DeviceBuffer":
        self = cls.__new__(cls)
        self.ptr, self.nbytes = ptr, nbytes
        return self

    def from_host(self, data: np.ndarray, offset: int = 0) -> None:
        data = np.ascontiguousarray(data)
        _check(lib().tpc_copy_to_device(C.c_void_p(self.ptr + offset), data.ctypes.data, data.nbytes))

    def to_host(self, nbytes: int | None = None, offset: int = 0) -> np.ndarray:
        n = self.nbytes - offset if nbytes is None else nbytes
        out = np.empty(n, dtype=np.uint8)
        _check(lib().tpc_copy_to_host(out.ctypes.data, C.c_void_p(self.ptr + offset), n))
        return out

    def close(self) -> None:
        if getattr(self, "ptr", None):
            lib().tpc_device_free(C.c_void_p(self.ptr))
            self.ptr = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


class DeviceGenome:
    """Packed genome resident in HBM: codes / n_mask device buffers + host record table."""

    def __init__(self, codes: DeviceBuffer, n_mask: DeviceBuffer, n_positions: int, rec_start: np.ndarray,
                 rec_len: np.ndarray, ascii_buf: DeviceBuffer | None = None):
        self.codes, self.n_mask, self.n_positions = codes, n_mask, int(n_positions)
        self.rec_start, self.rec_len, self.ascii = rec_start, rec_len, ascii_buf

    @property
    def total_bp(self) -> int:
        return int(self.rec_len.sum())

    def attach(self, session: "Session") -> None:
        session.set_genome_device(self.codes.ptr, self.n_mask.ptr, self.n_positions, self.rec_start, self.rec_len, keep=self)

    def to_host(self) -> PackedGenome:
        L = lib()
        cw, mw = L.tpc_code_words(self.n_positions), L.tpc_mask_words(self.n_positions)
        return PackedGenome(self.codes.to_host(cw * 8).view(np.uint64), self.n_mask.to_host(mw * 8).view(np.uint64),
                            self.n_positions, self.rec_start, self.rec_len)

    def record_ascii(self, r: int, length: int | None = None) -> bytes:
        """Bases of record r (prefix of `length`) copied back from the device ASCII buffer."""
        n = int(self.rec_len[r]) if length is None else min(int(length), int(self.rec_len[r]))
        return self.ascii.to_host(n, int(self.rec_start[r])).tobytes()


def pack_ascii_device(ascii_buf: DeviceBuffer, n_positions: int, rec_start: np.ndarray, rec_len: np.ndarray,
                      keep_ascii: bool = True) -> DeviceGenome:
    """K0 on the device: ASCII (tpc_genome layout) -> 2-bit codes + N mask."""
    L = lib()
    codes = DeviceBuffer(L.tpc_code_words(n_positions) * 8)
    n_mask = DeviceBuffer(L.tpc_mask_words(n_positions) * 8)
    _check(L.tpc_pack_ascii_device(C.c_void_p(ascii_buf.ptr), n_positions, C.c_void_p(codes.ptr), C.c_void_p(n_mask.ptr), None))
    g = DeviceGenome(codes, n_mask, n_positions, rec_start, rec_len, ascii_buf if keep_ascii else None)
    if not keep_ascii:
        ascii_buf.close()
    return g


class _DevArray:
    def __init__(self, ptr, n, typestr):
        self.__cuda_array_interface__ = {"shape": (n,), "typestr": typestr, "data": (ptr, False), "version": 2}


def as_torch(ptr: int, n: int, dtype):
    """Zero-copy torch view of a device pointer owned by the library (for torch.distributed)."""
    import torch
    typestr = {torch.int64: "<i8", torch.int32: "<i4", torch.uint8: "|u1"}[dtype]
    if n == 0 or not ptr:
        return torch.empty(0, dtype=dtype, device="cuda")
    return torch.as_tensor(_DevArray(ptr, n, typestr), device="cuda")
