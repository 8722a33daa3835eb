#!/usr/bin/env python
"""Registers / spills / shared memory per kernel from an `nvcc -Xptxas -v` log:  python tools/ptxas_regs.py LOG [filter]"""
import re
import subprocess
import sys

log = open(sys.argv[1]).read().splitlines()
flt = sys.argv[2] if len(sys.argv) > 2 else ""
rows, cur, spill = [], None, ""
for line in log:
    m = re.search(r"Compiling entry function '(\S+)'", line)
    if m:
        cur, spill = m.group(1), ""
    m = re.search(r"(\d+) bytes stack frame, (\d+) bytes spill stores, (\d+) bytes spill loads", line)
    if m and cur:
        spill = f"stack {m.group(1)} spill {m.group(2)}/{m.group(3)}"
    m = re.search(r"Used (\d+) registers(.*)", line)
    if m and cur:
        sm = re.search(r"(\d+) bytes smem", m.group(2))
        rows.append((cur, int(m.group(1)), sm.group(1) if sm else "0", spill))
        cur = None
names = subprocess.run(["c++filt"] + [r[0] for r in rows], capture_output=True, text=True).stdout.splitlines()
for nm, (_, regs, smem, spill) in zip(names, rows):
    nm = nm.replace("sg::", "").split("(")[0].replace("void ", "")
    if flt in nm:
        print(f"{regs:4d} regs  {smem:>6s} smem  {spill:28s} {nm[:90]}")
