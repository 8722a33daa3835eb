#!/usr/bin/env python
"""Dynamic (executed) instruction mix of one kernel from `ncu --page source --csv` output (optionally .gz):
warp-level executed instructions per opcode, per warp of the launch, and the stall samples each opcode class collects.

    tools/ncu_dynamic_mix.py gpurun_out/x_hot_kernel_source.csv.gz [warps]
"""
import collections
import csv
import gzip
import io
import sys


def main():
    path = sys.argv[1]
    f = io.TextIOWrapper(gzip.open(path)) if path.endswith(".gz") else open(path)
    rows = list(csv.reader(f))
    hdr = rows[1]
    ix = {h: i for i, h in enumerate(hdr)}
    ex, sm = collections.Counter(), collections.Counter()
    warps = None
    for r in rows[2:]:
        if len(r) < len(hdr):
            continue
        src = r[ix["Source"]].split()
        if not src:
            continue
        op = src[1] if src[0].startswith("@") else src[0]
        n = int(float(r[ix["Instructions Executed"]] or 0))
        if warps is None:
            warps = n
        base = op.split(".")[0]
        if base == "MUFU" or op.startswith("IMAD.MOV") or op.startswith("IMAD.WIDE"):
            base = ".".join(op.split(".")[:2])
        ex[base] += n
        sm[base] += int(float(r[ix["# Samples"]] or 0))
    if len(sys.argv) > 2:
        warps = int(sys.argv[2])
    tot = sum(ex.values())
    fp64 = sum(v for k, v in ex.items() if k in ("DFMA", "DMUL", "DADD", "DSETP") or k.startswith("MUFU.RCP64") or k.startswith("MUFU.RSQ64"))
    print(f"warps {warps}  executed {tot}  per warp {tot / warps:.1f}  fp64-pipe per warp {fp64 / warps:.1f} ({100 * fp64 / tot:.1f} %)")
    print(f"{'opcode':14s} {'per warp':>9s} {'share':>7s} {'samples':>8s}")
    for k, v in ex.most_common(40):
        print(f"{k:14s} {v / warps:9.1f} {100 * v / tot:6.1f}% {sm[k]:8d}")


if __name__ == "__main__":
    main()
