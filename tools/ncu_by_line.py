#!/usr/bin/env python
"""Executed instructions and stall samples of one kernel per CUDA SOURCE LINE (innermost inlined frame).

    ncu -i rep.ncu-rep --page source --csv > sass.csv                       # per-SASS-instruction counters
    cuobjdump -xelf all lib.so; nvdisasm -gi x.cubin > dis.txt               # SASS annotated with line info
    tools/ncu_by_line.py sass.csv dis.txt <mangled kernel name> [top]

Joins the two by instruction order.  Prints the top lines by executed warp instructions (with the share that goes to
the FP64 pipe and the moves among them) and the per-function totals (function = nearest preceding line in the same file
that looks like a definition)."""
import collections
import csv
import re
import sys


def main():
    sass_csv, dis, kernel = sys.argv[1:4]
    top = int(sys.argv[4]) if len(sys.argv) > 4 else 40
    rows = list(csv.reader(open(sass_csv)))
    hdr = rows[1]
    ix = {h: i for i, h in enumerate(hdr)}
    data = [r for r in rows[2:] if len(r) >= len(hdr)]
    # annotated disassembly of the kernel
    lines = open(dis).read().splitlines()
    start = next(i for i, l in enumerate(lines) if l.startswith(".text." + kernel + ":"))
    ann, cur, fresh = [], None, True
    for l in lines[start + 1:]:
        if l.startswith(".text.") or l.startswith("//-----"):
            break
        m = re.match(r'\s*//## File "([^"]+)", line (\d+)', l)
        if m:
            if fresh:  # the first annotation of a group is the innermost frame
                cur = (m.group(1).split("/")[-1], int(m.group(2)))
                fresh = False
            continue
        m = re.match(r"\s*/\*([0-9a-f]+)\*/\s+(.*?);", l)
        if m:
            ann.append((cur, m.group(2).strip()))
            fresh = True
    assert len(ann) == len(data), (len(ann), len(data))
    ex, smp, fp, mv = collections.Counter(), collections.Counter(), collections.Counter(), collections.Counter()
    warps = int(float(data[0][ix["Instructions Executed"]]))
    for (loc, text), r in zip(ann, data):
        n = int(float(r[ix["Instructions Executed"]] or 0))
        s = int(float(r[ix["# Samples"]] or 0))
        op = text.split()[1] if text.startswith("@") else text.split()[0]
        ex[loc] += n
        smp[loc] += s
        if op.split(".")[0] in ("DFMA", "DMUL", "DADD", "DSETP") or op.startswith("MUFU.RCP64") or op.startswith("MUFU.RSQ64"):
            fp[loc] += n
        if op.startswith("IMAD.MOV") or op.split(".")[0] in ("MOV", "UMOV", "CS2R", "FSEL", "SEL"):
            mv[loc] += n
    src = {}
    for f in ("sg_engine.cu", "sg_device.cuh", "sg_tail.cuh"):
        try:
            src[f] = open("spacegym_b200/csrc/" + f).read().splitlines()
        except OSError:
            pass

    def func_of(loc):
        f, ln = loc
        if f not in src:
            return f
        for k in range(min(ln, len(src[f])) - 1, -1, -1):
            m = re.match(r"^(?:static\s+)?(?:__device__|__global__).*?(\w+)\s*\(", src[f][k])
            if m and not src[f][k].startswith(" "):
                return m.group(1)
            m = re.match(r"^(\w+)\(", src[f][k])  # name on its own line after the qualifiers
            if m:
                return m.group(1)
        return f

    tot = sum(ex.values())
    print(f"warps {warps}, executed per warp {tot / warps:.1f}, samples {sum(smp.values())}")
    fn_ex, fn_smp, fn_fp, fn_mv = collections.Counter(), collections.Counter(), collections.Counter(), collections.Counter()
    for loc in ex:
        fn = func_of(loc) if loc else "?"
        fn_ex[fn] += ex[loc]; fn_smp[fn] += smp[loc]; fn_fp[fn] += fp[loc]; fn_mv[fn] += mv[loc]
    print(f"{'function':28s} {'instr/warp':>10s} {'fp64':>7s} {'moves/sel':>9s} {'samples':>8s}")
    for fn, n in fn_ex.most_common(30):
        print(f"{fn:28s} {n / warps:10.1f} {fn_fp[fn] / warps:7.1f} {fn_mv[fn] / warps:9.1f} {fn_smp[fn]:8d}")
    print()
    print(f"{'line':22s} {'instr/warp':>10s} {'fp64':>6s} {'mov':>6s} {'smp':>5s}  source")
    for loc, n in ex.most_common(top):
        f, ln = loc if loc else ("?", 0)
        text = src.get(f, [""] * (ln + 1))[ln - 1].strip()[:90] if f in src and ln <= len(src[f]) else ""
        print(f"{f + ':' + str(ln):22s} {n / warps:10.1f} {fp[loc] / warps:6.1f} {mv[loc] / warps:6.1f} {smp[loc]:5d}  {text}")


if __name__ == "__main__":
    main()
