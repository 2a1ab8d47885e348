#!/usr/bin/env python
"""Opcode counts and the memory / synchronisation instructions of the hot kernels, from `cuobjdump -sass` of the built objects
(no GPU needed).   python tools/sass_excerpt.py > profiles/r2_sass_excerpt.md"""
import collections
import re
import subprocess
import sys
from pathlib import Path

BUILD = Path(__file__).resolve().parents[1] / "twopaco_b200" / "csrc" / "build"
KERNELS = [  # (object, regex on the mangled name)
    ("tpc_session.o", r"k_apply_queryILi5ELb1"), ("tpc_session.o", r"k_apply_fillILi5"),
    ("tpc_w1.o", r"k_bin_listILi1ELi8ELb0"), ("tpc_w1.o", r"k_bin_listILi1ELi8ELb1"), ("tpc_w1.o", r"k_ownILi1ELi3"), ("tpc_w1.o", r"5k_binILi1"),
    ("tpc_w1.o", r"k_emit_writeILi1"), ("tpc_w1.o", r"k_emit_countILi1"), ("tpc_w1.o", r"8k_insertILi1"), ("tpc_w1.o", r"k_insert_listILi1"),
    ("tpc_w1.o", r"k_direct_listILi1ELi5ELb1"), ("tpc_wn19.o", r"k_bin_listILi19ELi8ELb0"),
]
MEM = re.compile(r"^(LDG|STG|LDS|STS|LDGSTS|ATOM|RED|BAR|SHFL|VOTE|MATCH|WARPSYNC|UTMA|UBLKCP|SYNCS|LDGDEPBAR|DEPBAR)")

print("# SASS of the hot kernels of libtwopaco_b200.so (cuobjdump -sass, sm_100a, final round-2 build; tools/sass_excerpt.py): opcode counts and")
print("# the memory / synchronisation instructions.  LDG.E...256 = one 256-bit load per 32-byte filter sector (sm_100); LDGSTS = cp.async;")
print("# REDG / ATOMG = atomicOr / CAS on the filter, the mask and the tables; VOTE + POPC = the ballot slots of the mark queue;")
print("# no UTMALDG / UBLKCP (TMA) and no UTC*MMA (tcgen05): see DESIGN.md 'Which B200 features the path uses'\n")
for obj, pat in KERNELS:
    names = subprocess.run(["cuobjdump", "-elf", str(BUILD / obj)], capture_output=True, text=True).stdout
    m = re.search(r"\.text\.(\S*" + pat + r"\S*)", names)
    if not m:
        print(f"## ({pat}: not found in {obj})\n")
        continue
    name = m.group(1)
    sass = subprocess.run(["cuobjdump", "-sass", "-fun", name, str(BUILD / obj)], capture_output=True, text=True).stdout
    ops = [l.split()[1].rstrip(";") for l in sass.splitlines() if re.match(r"^\s+/\*[0-9a-f]{4}\*/", l) and len(l.split()) > 1]
    ops = [o if not o.startswith("@") else None for o in ops]
    # predicated instructions: the opcode is the next token
    full = []
    for l in sass.splitlines():
        if not re.match(r"^\s+/\*[0-9a-f]{4}\*/", l):
            continue
        t = l.split()[1:]
        if t and t[0].startswith("@"):
            t = t[1:]
        if t:
            full.append(t[0].rstrip(";"))
    base = collections.Counter(o.split(".")[0] for o in full)
    mem = collections.Counter(o for o in full if MEM.match(o))
    demangled = subprocess.run(["c++filt", name], capture_output=True, text=True).stdout.strip().split("(")[0]
    print(f"## {demangled}")
    print(f"{len(full)} instructions; top opcodes: " + ", ".join(f"{k} {v}" for k, v in base.most_common(10)))
    print("memory / sync: " + ", ".join(f"{k} x{v}" for k, v in sorted(mem.items())) + "\n")
