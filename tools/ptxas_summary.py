#!/usr/bin/env python
"""Summarise `nvcc -Xptxas -v` logs: kernel, registers, stack, spills, shared memory."""
import re, subprocess, sys
for path in sys.argv[1:]:
    txt = open(path).read()
    for m in re.finditer(r"Compiling entry function '(\S+)' for 'sm_100a'\s*\nptxas info\s*: Function properties for \S+\s*\n\s*(\d+) bytes stack frame, (\d+) bytes spill stores, (\d+) bytes spill loads\s*\nptxas info\s*: Used (\d+) registers(.*)", txt):
        name = subprocess.run(["c++filt", m.group(1)], capture_output=True, text=True).stdout.strip()
        name = re.sub(r"\(.*", "", name).replace("void tpc::", "")
        smem = re.search(r"(\d+) bytes smem", m.group(6))
        print(f"{name:32s} regs={m.group(5):>3s} stack={m.group(2):>3s} spill={m.group(3)}/{m.group(4)} smem={smem.group(1) if smem else 0}")
