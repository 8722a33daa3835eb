#!/usr/bin/env python
"""Turn the files a profiling `gpurun` call left in gpurun_out/ into the tracked summaries under profiles/.

    tools/collect_profiles.py <prefix> <tag>      e.g.  tools/collect_profiles.py s3i r1_d

expects gpurun_out/<prefix>_{hot,cont,tail}_kernel_raw.csv (`ncu -i rep --page raw --csv` of one `--set full`
capture per kernel), <prefix>_{...}_summary.txt (tools/ncu_summary.py), <prefix>_launches.csv (launch list),
<prefix>_bench.json and optionally <prefix>_ref.json; writes profiles/<tag>_ncu_final_summary.csv,
<tag>_*_stalls.txt, <tag>_launches_final.csv, <tag>_bench_final.json, <tag>_bench_reference_arm.json and
traffic.json (DRAM bytes of one launch of the dominant kernel, read by bench.py for `roofline.traffic`).
"""
import collections
import csv
import json
import os
import shutil
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
COLS = [
    "gpu__time_duration.sum", "launch__grid_size", "launch__block_size", "launch__registers_per_thread",
    "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem",
    "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__inst_executed.sum",
    "smsp__thread_inst_executed_per_inst_executed.ratio", "smsp__issue_active.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
    "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
    "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_wait_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_not_selected_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_no_instruction_per_issue_active.ratio",
    "smsp__warps_eligible.avg.per_cycle_active",
]
KERNELS = ("hot_kernel", "cont_kernel", "tail2_kernel")  # round 1: tail_kernel


def main():
    prefix, tag = sys.argv[1], sys.argv[2]
    src = lambda name: os.path.join(ROOT, "gpurun_out", f"{prefix}_{name}")  # noqa: E731
    dst = lambda name: os.path.join(ROOT, "profiles", f"{tag}_{name}")  # noqa: E731
    out, units, traffic = [], None, {}
    for k in KERNELS:
        rows = list(csv.reader(open(src(f"{k}_raw.csv"))))
        d, u = dict(zip(rows[0], rows[2])), dict(zip(rows[0], rows[1]))
        if units is None:
            units = [""] + [u.get(c, "") for c in COLS]
        out.append([d["Kernel Name"][:70]] + [d.get(c, "") for c in COLS])
        scale = {"byte": 1, "kbyte": 1e3, "mbyte": 1e6, "gbyte": 1e9}
        traffic[k] = int(sum(float(d[m]) * scale[u[m].lower()] for m in ("dram__bytes_read.sum", "dram__bytes_write.sum")))
        shutil.copy(src(f"{k}_summary.txt"), dst(f"{k}_stalls.txt"))
        if os.path.exists(src(f"{k}_source.csv.gz")):  # dynamic instruction mix (tools/ncu_dynamic_mix.py)
            import subprocess
            mix = subprocess.run([sys.executable, os.path.join(ROOT, "tools", "ncu_dynamic_mix.py"), src(f"{k}_source.csv.gz")],
                                 capture_output=True, text=True).stdout
            open(dst(f"{k}_dynamic_mix.txt"), "w").write(mix)
        print(k, d["gpu__time_duration.sum"], "us", d["launch__registers_per_thread"], "regs, issue",
              d["smsp__issue_active.avg.pct_of_peak_sustained_active"], "fp64",
              d["sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active"], "inst", d["smsp__inst_executed.sum"])
    with open(dst("ncu_final_summary.csv"), "w", newline="") as f:
        w = csv.writer(f)
        w.writerow(["Kernel Name"] + COLS)
        w.writerow(units)
        w.writerows(out)
    json.dump({"dram_bytes_per_launch": traffic["hot_kernel"],
               "what": "dram__bytes_read.sum + dram__bytes_write.sum of ONE launch of the dominant kernel (hot_kernel, "
                       f"262144 envs), ncu --set full, profiles/{tag}_ncu_final_summary.csv; reads match the 136 B/env of "
                       "state + action (35.7 MB), most of the kernel's 50 MB of writes are still in the write-back L2 "
                       "when it ends",
               "per_kernel": traffic, "algorithmic_bytes_hot_kernel": int(329 * 193121),
               "algorithmic_bytes_whole_step": 86245376},
              open(os.path.join(ROOT, "profiles", "traffic.json"), "w"), indent=1)
    # executed FP64-pipe warp instructions per kernel (bench.py: roofline_fp64_pipe_whole_step)
    import re
    fp = {}
    for k in KERNELS:
        if os.path.exists(dst(f"{k}_dynamic_mix.txt")):
            m = re.match(r"warps (\d+)\s+executed (\d+)\s+per warp ([\d.]+)\s+fp64-pipe per warp ([\d.]+)",
                         open(dst(f"{k}_dynamic_mix.txt")).readline())
            if m:
                fp[k] = {"warps": int(m.group(1)), "executed_warp_instructions": int(m.group(2)),
                         "fp64_pipe_warp_instructions": round(float(m.group(4)) * int(m.group(1)))}
    if len(fp) == len(KERNELS):
        tj = json.load(open(os.path.join(ROOT, "profiles", "traffic.json")))
        tj["fp64_pipe"] = {"per_kernel": fp,
                           "fp64_pipe_warp_instructions_per_step": sum(v["fp64_pipe_warp_instructions"] for v in fp.values()),
                           "what": "executed FP64-pipe warp instructions (DFMA, DMUL, DADD, DSETP, MUFU.RCP64H / RSQ64H) of the "
                                   f"three kernels of ONE 262144-env step, ncu source-page counters of profiles/{tag}_* "
                                   "(tools/ncu_dynamic_mix.py); the pipe issues 0.5 warp instructions per cycle per scheduler"}
        json.dump(tj, open(os.path.join(ROOT, "profiles", "traffic.json"), "w"), indent=1)
    shutil.copy(src("launches.csv"), dst("launches_final.csv"))
    shutil.copy(src("bench.json"), dst("bench_final.json"))
    if os.path.exists(src("ref.json")):
        shutil.copy(src("ref.json"), dst("bench_reference_arm.json"))
    agg = collections.defaultdict(list)
    for r in csv.reader(open(dst("launches_final.csv"))):
        if len(r) > 10 and r[0].isdigit():
            agg[r[4].split("<")[0].replace("void ", "").replace("sg::", "")].append(float(r[-1]) / 1000)
    mean = {k: sum(agg[k]) / len(agg[k]) for k in KERNELS}
    tot = sum(mean.values())
    print("launch list (us):", {k: round(v, 1) for k, v in mean.items()}, "shares:",
          {k: round(v / tot, 3) for k, v in mean.items()})
    b = json.load(open(dst("bench_final.json")))
    kt = b["kernel_times"]
    ev = {"hot_kernel": kt["hot_us"], "cont_kernel": kt["cont_us"], KERNELS[2]: kt["tail_us"]}
    print("CUDA events (us):", {k: round(v, 1) for k, v in ev.items()}, "shares:",
          {k: round(v / sum(ev.values()), 3) for k, v in ev.items()})
    print("value", b["value"], "e2e", b["e2e"]["value"], "roofline", b["roofline"]["frac"], b["roofline_fp64"]["frac"],
          "whole", b["roofline_whole_step"]["hbm"]["frac"], b["roofline_whole_step"]["fp64"]["frac"])


if __name__ == "__main__":
    main()
