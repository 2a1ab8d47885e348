"""ctypes binding of tools/benchutil/libtpc_benchutil.so -- BENCH / TEST INFRASTRUCTURE ONLY.

Synthetic founder-family inputs generated on the device and the roofline probes
(random sectors in HBM; random sectors inside one L2-resident filter slice).  None of this is part of
the product ABI (include/twopaco_b200.h) and nothing under twopaco_b200/ imports it.
"""
from __future__ import annotations

import ctypes as C
from pathlib import Path

import numpy as np

from twopaco_b200 import api

LIB_PATH = Path(__file__).resolve().parent / "libtpc_benchutil.so"
_lib = None


class BenchUtilError(RuntimeError):
    pass


def lib() -> C.CDLL:
    global _lib
    if _lib is None:
        if not LIB_PATH.exists():
            raise BenchUtilError(f"{LIB_PATH} is missing: run `python -c 'import __graft_entry__ as g; g.build()'`")
        L = C.CDLL(str(LIB_PATH))
        L.tpcb_last_error.restype = C.c_char_p
        L.tpcb_synth_family_device.restype = C.c_int
        L.tpcb_synth_family_device.argtypes = [C.c_uint64, C.c_uint32, C.c_uint32, C.c_uint64, C.c_double, C.POINTER(C.c_void_p),
                                               C.POINTER(C.c_uint64), C.c_void_p, C.c_void_p]
        L.tpcb_synth_family_part_device.restype = C.c_int
        L.tpcb_synth_family_part_device.argtypes = [C.c_uint64, C.c_uint32, C.c_uint32, C.c_uint32, C.c_uint64, C.c_double, C.POINTER(C.c_void_p),
                                                    C.POINTER(C.c_uint64), C.c_void_p, C.c_void_p]
        L.tpcb_random_access_probe.restype = C.c_int
        L.tpcb_random_access_probe.argtypes = [C.c_uint32, C.c_uint32, C.c_uint64, C.POINTER(C.c_double)]
        L.tpcb_slice_probe.restype = C.c_int
        L.tpcb_slice_probe.argtypes = [C.c_uint32, C.c_uint32, C.c_uint64, C.c_uint32, C.c_uint32, C.c_uint32, C.c_uint32,
                                       C.POINTER(C.c_double)]
        L.tpcb_rank_probe.restype = C.c_int
        L.tpcb_rank_probe.argtypes = [C.c_uint32, C.c_uint32, C.c_uint32, C.POINTER(C.c_double)]
        _lib = L
    return _lib


def _check(rc: int) -> None:
    if rc != 0:
        raise BenchUtilError(lib().tpcb_last_error().decode(errors="replace"))


def synth_family_device(seed: int, genomes: int, records_per_genome: int, record_len: int, p: float,
                        keep_ascii: bool = True) -> api.DeviceGenome:
    """Founder-family genome set generated on the device, then packed by the product's K0."""
    L = lib()
    n = genomes * records_per_genome
    rec_start = np.empty(n, dtype=np.uint64)
    rec_len = np.empty(n, dtype=np.uint64)
    ptr, npos = C.c_void_p(), C.c_uint64()
    _check(L.tpcb_synth_family_device(seed, genomes, records_per_genome, record_len, p, C.byref(ptr), C.byref(npos),
                                      rec_start.ctypes.data, rec_len.ctypes.data))
    buf = api.DeviceBuffer.adopt(ptr.value, (npos.value + 63) // 64 * 64 + 64)
    return api.pack_ascii_device(buf, npos.value, rec_start, rec_len, keep_ascii=keep_ascii)


def random_access_probe(filter_bits: int, mode: int, touches: int = 1 << 30) -> float:
    """Uniform random 32-byte sector touches per second into a 2^filter_bits-bit table in HBM.
    mode 0 = 32-byte loads, 1 = atomicOr, 2 = load + conditional atomicOr."""
    v = C.c_double()
    _check(lib().tpcb_random_access_probe(filter_bits, mode, touches, C.byref(v)))
    return v.value


def rank_probe(mode: int, buckets: int = 128, rounds: int = 2000) -> float:
    """Records ranked per second by the inner step of a CTA-level counting sort (see tpcb_rank_probe)."""
    v = C.c_double()
    _check(lib().tpcb_rank_probe(mode, buckets, rounds, C.byref(v)))
    return v.value


def slice_probe(slice_log2: int = 26, slices: int = 8, records_per_slice: int = 32 << 20, dup: int = 7, mode: int = 0,
                U: int = 4, ctas_per_sm: int = 4) -> float:
    """Random sector touches per second inside one L2-resident filter slice with the 8-byte record stream
    read beside it (the pattern of k_apply_query / k_apply_fill).  mode 0 = load, 1 = query test, 2 = fill."""
    v = C.c_double()
    _check(lib().tpcb_slice_probe(slice_log2, slices, records_per_slice, dup, mode, U, ctas_per_sm, C.byref(v)))
    return v.value


def synth_family_host(seed: int, genomes: int, records_per_genome: int, record_len: int, p: float, group: int = 8,
                      pin: bool = True) -> api.PackedGenome:
    """The founder family generated on the device in groups of `group` genomes and packed straight into (pinned) HOST
    memory: for sets whose packed form does not fit HBM (BASELINE config 5: 100 haplotypes, 310 Gbp, 116 GB packed).
    Same bases as synth_family_device.  Every group is packed by the product's K0 at its own bit offset: the last < 64
    bases of the stream so far are carried over so that the group's first word is complete."""
    import torch
    L, P = lib(), api.lib()
    n_rec = genomes * records_per_genome
    # upper bound of the positions (insertions lengthen a record by ~ p / 10)
    max_pos = 1 + n_rec * (int(record_len * (1 + p)) + 64)
    cw_cap, mw_cap = P.tpc_code_words(max_pos), P.tpc_mask_words(max_pos)
    # (pinned memory is allocated as such: pinning a pageable copy would need the 116 GB of C5 twice)
    mk = (lambda n: torch.empty(n, dtype=torch.int64, pin_memory=True)) if pin else (lambda n: torch.empty(n, dtype=torch.int64))
    codes, nmask = mk(cw_cap), mk(mw_cap)
    codes_np, nmask_np = codes.numpy().view(np.uint64), nmask.numpy().view(np.uint64)
    rec_start = np.empty(n_rec, dtype=np.uint64)
    rec_len = np.empty(n_rec, dtype=np.uint64)
    pos = 1                                            # global position of the next base; position 0 is the leading 'N'
    carry = torch.full((1,), ord("N"), dtype=torch.uint8, device="cuda")   # ASCII of the positions [pos - pos % 64, pos)
    for g0 in range(0, genomes, group):
        ng = min(group, genomes - g0)
        nr = ng * records_per_genome
        rs, rl = np.empty(nr, dtype=np.uint64), np.empty(nr, dtype=np.uint64)
        ptr, npos = C.c_void_p(), C.c_uint64()
        _check(L.tpcb_synth_family_part_device(seed, g0, ng, records_per_genome, record_len, p, C.byref(ptr), C.byref(npos),
                                               rs.ctypes.data, rl.ctypes.data))
        n = npos.value - 1                             # bytes of this group after its own leading 'N'
        a = carry.numel()                              # == pos % 64 (1 for the very first group: the leading 'N')
        ascii_buf = api.DeviceBuffer.adopt(ptr.value, (npos.value + 63) // 64 * 64 + 64)
        src = api.as_torch(ptr.value + 1, n, torch.uint8)
        tmp_bytes = (a + n + 63) // 64 * 64 + 128
        tmp = torch.full((tmp_bytes,), ord("N"), dtype=torch.uint8, device="cuda")
        tmp[:a].copy_(carry)
        tmp[a:a + n].copy_(src)
        torch.cuda.synchronize()
        ascii_buf.close()
        base = pos - a                                 # global position of tmp[0]: a multiple of 64
        assert base % 64 == 0
        tmp_buf = api.DeviceBuffer.adopt(tmp.data_ptr(), tmp_bytes)
        dg = api.pack_ascii_device(tmp_buf, a + n, np.zeros(0, np.uint64), np.zeros(0, np.uint64), keep_ascii=True)
        tmp_buf.ptr = None                             # (memory owned by the torch tensor)
        cw, mw = (a + n + 31) // 32, (a + n + 63) // 64
        api._check(P.tpc_copy_to_host(codes_np[base // 32:].ctypes.data, C.c_void_p(dg.codes.ptr), cw * 8))
        api._check(P.tpc_copy_to_host(nmask_np[base // 64:].ctypes.data, C.c_void_p(dg.n_mask.ptr), mw * 8))
        dg.codes.close(); dg.n_mask.close()
        rec_start[g0 * records_per_genome:g0 * records_per_genome + nr] = rs - 1 + pos
        rec_len[g0 * records_per_genome:g0 * records_per_genome + nr] = rl
        pos += n
        carry = tmp[a + n - (pos % 64):a + n].clone() if pos % 64 else torch.empty(0, dtype=torch.uint8, device="cuda")
        del tmp, src
    n_positions = pos
    cw, mw = P.tpc_code_words(n_positions), P.tpc_mask_words(n_positions)
    # padding after the last position: codes 0, n-mask 1 (the last group's pack already padded its final words)
    last_c, last_m = (n_positions + 31) // 32, (n_positions + 63) // 64
    codes_np[last_c:cw] = 0
    nmask_np[last_m:mw] = np.uint64(0xFFFFFFFFFFFFFFFF)
    g = api.PackedGenome(codes_np[:cw], nmask_np[:mw], n_positions, rec_start, rec_len)
    g._keep = (codes, nmask)
    return g
