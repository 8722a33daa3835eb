#!/usr/bin/env python
"""Static SASS instruction mix of the kernels of a built engine library (no GPU needed).

    python tools/sass_hist.py spacegym_b200/libspacegym_b200.so hot_kernel [--top 30] [--dump out.sass]

Prints, per matching kernel (substring of the mangled name), the instruction count by opcode, the share of FP64-pipe
instructions (DFMA/DMUL/DADD/DSETP/MUFU.*64), and the count of instructions that only materialise constants or move
registers (IMAD.MOV / MOV / UMOV / CS2R) — the number VERDICT r1 asks to watch.
"""
import collections
import re
import subprocess
import sys


def functions(lib):
    out = subprocess.run(["cuobjdump", "-sass", lib], capture_output=True, text=True, check=True).stdout
    name, rows = None, []
    for line in out.splitlines():
        m = re.match(r"\s+Function : (\S+)", line)
        if m:
            if name:
                yield name, rows
            name, rows = m.group(1), []
            continue
        m = re.match(r"\s+/\*([0-9a-f]{4,})\*/\s+(.*?)\s*;?\s*/\*", line)
        if m and name:
            rows.append(m.group(2))
    if name:
        yield name, rows


def main():
    lib, pat = sys.argv[1], sys.argv[2]
    top = int(sys.argv[sys.argv.index("--top") + 1]) if "--top" in sys.argv else 25
    dump = sys.argv[sys.argv.index("--dump") + 1] if "--dump" in sys.argv else None
    for name, rows in functions(lib):
        if pat not in name:
            continue
        ops = collections.Counter()
        movs = fp64 = 0
        for r in rows:
            r = re.sub(r"^@!?U?P\w+\s+", "", r)
            op = r.split()[0].rstrip(";")
            base = op.split(".")[0]
            ops[base] += 1
            if op.startswith(("IMAD.MOV", "MOV", "UMOV", "CS2R")) or (base == "IMAD" and "RZ, RZ" in r):
                movs += 1
            if base in ("DFMA", "DMUL", "DADD", "DSETP") or (base == "MUFU" and "64" in op):
                fp64 += 1
        n = len(rows)
        print(f"{name[:110]}\n  instructions {n}  fp64-pipe {fp64} ({100.0 * fp64 / max(n, 1):.1f} %)  "
              f"moves/constants {movs} ({100.0 * movs / max(n, 1):.1f} %)")
        print("  " + "  ".join(f"{k} {v}" for k, v in ops.most_common(top)))
        if dump:
            with open(dump, "w") as f:
                f.write("\n".join(rows) + "\n")


if __name__ == "__main__":
    main()
