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
        L.tpcb_random_access_probe.restype = C.c_int
        L.tpcb_random_access_probe.argtypes = [C.c_uint32, C.c_uint32, C.c_uint64, C.POINTER(C.c_double)]
        L.tpcb_slice_probe.restype = C.c_int
        L.tpcb_slice_probe.argtypes = [C.c_uint32, C.c_uint32, C.c_uint64, C.c_uint32, C.c_uint32, C.c_uint32, C.c_uint32,
                                       C.POINTER(C.c_double)]
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


def slice_probe(slice_log2: int = 26, slices: int = 8, records_per_slice: int = 32 << 20, dup: int = 7, mode: int = 0,
                U: int = 4, ctas_per_sm: int = 4) -> float:
    """Random sector touches per second inside one L2-resident filter slice with the 8-byte record stream
    read beside it (the pattern of k_apply_query / k_apply_fill).  mode 0 = load, 1 = query test, 2 = fill."""
    v = C.c_double()
    _check(lib().tpcb_slice_probe(slice_log2, slices, records_per_slice, dup, mode, U, ctas_per_sm, C.byref(v)))
    return v.value
