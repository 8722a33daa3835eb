#!/usr/bin/env python
"""Summarise an .ncu-rep (one kernel): headline metrics, stall breakdown per issue, stall samples per
100-instruction block and the top stalled SASS instructions.  Usage: tools/ncu_summary.py report.ncu-rep [top]"""
import csv
import io
import subprocess
import sys


def page(rep, name):
    out = subprocess.run(["ncu", "-i", rep, "--page", name, "--csv"], capture_output=True, text=True).stdout
    return list(csv.reader(io.StringIO(out)))


def main():
    rep = sys.argv[1]
    top_n = int(sys.argv[2]) if len(sys.argv) > 2 else 25
    rows = page(rep, "raw")
    d = dict(zip(rows[0], rows[2]))
    print("kernel:", d.get("Kernel Name", "")[:90])
    for k in ("gpu__time_duration.sum", "launch__grid_size", "launch__block_size", "launch__registers_per_thread",
              "smsp__inst_executed.sum", "smsp__thread_inst_executed_per_inst_executed.ratio",
              "smsp__issue_active.avg.pct_of_peak_sustained_active",
              "sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active",
              "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__warps_eligible.avg.per_cycle_active",
              "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed"):
        print(f"  {k} = {d.get(k)}")
    print("stall cycles per issue:")
    st = {k.split("issue_stalled_")[1].split("_per_issue")[0]: float(v) for k, v in d.items()
          if k.startswith("smsp__average_warps_issue_stalled_") and k.endswith("_per_issue_active.ratio")}
    for k, v in sorted(st.items(), key=lambda kv: -kv[1]):
        if v > 0.005:
            print(f"  {k:22s} {v:6.3f}")
    rows = page(rep, "source")
    hdr, data = rows[1], rows[2:]
    ix = {h: i for i, h in enumerate(hdr)}

    def f(r, k):
        try:
            return float(r[ix[k]])
        except (ValueError, KeyError):
            return 0.0

    tot = sum(f(r, "# Samples") for r in data)
    print(f"samples {int(tot)}, per 100-instruction block: samples  exec/warp  lanes  long_sb short_sb wait")
    warps = max(f(r, "Instructions Executed") for r in data) or 1
    for b in range(0, len(data), 100):
        blk = data[b:b + 100]
        s = sum(f(r, "# Samples") for r in blk)
        if not s:
            continue
        ex = sum(f(r, "Instructions Executed") for r in blk)
        th = sum(f(r, "Thread Instructions Executed") for r in blk) / max(1.0, ex)
        print(f"  {b:5d} {int(s):6d} {ex / warps:8.1f} {th:6.1f}", *(f"{int(sum(f(r, k) for r in blk)):6d}"
              for k in ("stall_long_sb", "stall_short_sb", "stall_wait")))
    print("top stalled instructions: addr samples long_sb short_sb wait lanes exec source")
    for r in sorted(data, key=lambda r: -f(r, "# Samples"))[:top_n]:
        print(f"  {r[ix['Address']][-5:]} {int(f(r, '# Samples')):5d} {int(f(r, 'stall_long_sb')):5d} "
              f"{int(f(r, 'stall_short_sb')):5d} {int(f(r, 'stall_wait')):5d} {r[ix['Avg. Threads Executed']]:>4s} "
              f"{int(f(r, 'Instructions Executed')):7d}  {r[ix['Source']][:70]}")


if __name__ == "__main__":
    main()
