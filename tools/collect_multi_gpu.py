#!/usr/bin/env python
"""gpurun_out/mg{1,2,4,8}/ (tools/multi_gpu_round.sh) -> profiles/r2_mg*.json, r2_host_link_*gpu.json, r2_config5_sweep.json
and the markdown tables of DESIGN.md §4 on stdout."""
import json
import os
import shutil

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
G = [g for g in (1, 2, 4, 8) if os.path.isdir(os.path.join(ROOT, "gpurun_out", f"mg{g}"))]


def last_json(path):
    return json.loads(open(path).read().strip().splitlines()[-1])


def main():
    sweep, rows = [], {}
    for g in G:
        d = os.path.join(ROOT, "gpurun_out", f"mg{g}")
        shutil.copy(os.path.join(d, "host_link.json"), os.path.join(ROOT, "profiles", f"r2_host_link_{g}gpu.json"))
        shutil.copy(os.path.join(d, "topo.txt"), os.path.join(ROOT, "profiles", f"r2_topo_{g}gpu.txt"))
        for f in ("bench_s20", "bench_s3000"):
            p = os.path.join(d, f + ".json")
            if os.path.exists(p) and os.path.getsize(p):
                shutil.copy(p, os.path.join(ROOT, "profiles", f"r2_mg{g}_{f}.json"))
                rows.setdefault(g, {})[f] = last_json(p)
        link = last_json(os.path.join(d, "host_link.json"))[str(g)]
        rows.setdefault(g, {})["link"] = link
        for line in open(os.path.join(d, "sweep.jsonl")):
            sweep.append(json.loads(line))
    json.dump({"what": "BASELINE configs[4]: GoalContinuous4P-v0, one handle per GPU, tools/sweep_config5.py",
               "rows": sweep}, open(os.path.join(ROOT, "profiles", "r2_config5_sweep.json"), "w"), indent=1)
    print("| GPUs | --steps 20 | long run | e2e | link in the same run | e2e / link | plain link (host_link_peak.py) |")
    for g in G:
        r = rows[g]
        s20, lg = r.get("bench_s20"), r.get("bench_s3000")
        b = lg or s20
        print(f"| {g} | {s20['value']:.3g} | {lg['value']:.3g} | {b['e2e']['value']:.3g} | "
              f"{b['e2e']['link_gbs_measured']:.1f} | {b['e2e']['frac']:.2f} | {r['link']['d2h_gbs']:.1f} "
              f"({', '.join('%.1f' % x for x in r['link']['per_gpu_d2h_gbs'])}) |" if lg else
              f"| {g} | {s20['value']:.3g} | - | {s20['e2e']['value']:.3g} | {s20['e2e']['link_gbs_measured']:.1f} | "
              f"{s20['e2e']['frac']:.2f} | {r['link']['d2h_gbs']:.1f} |")
    totals = sorted({r["total_envs"] for r in sweep})
    print()
    print("| total envs | " + " | ".join(f"{g} GPU{'s' if g > 1 else ''}" for g in G) + " |")
    print("|---|" + "---|" * len(G))
    for t in totals:
        cells = []
        for g in G:
            m = [r for r in sweep if r["total_envs"] == t and r["n_gpus"] == g]
            cells.append(f"{m[0]['value']:.3g} ({m[0]['e2e']:.2g})" if m else "—")
        print(f"| 2^{t.bit_length() - 1} = {t} | " + " | ".join(cells) + " |")


if __name__ == "__main__":
    main()
