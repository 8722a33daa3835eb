#!/usr/bin/env python
"""Host-link roofline of the end-to-end (host buffer) path: plain pinned cudaMemcpyAsync, nothing else.

    python tools/host_link_peak.py [--gpus 1,2,4,8] [--mb 34] [--seconds 1.0] [--out profiles/r2_host_link.json]

For every GPU count G in --gpus (capped at the visible devices) one worker PROCESS per GPU (what bench.py is under
torchrun) copies pinned host buffers of --mb megabytes, one cudaMemcpyAsync per buffer on its own stream, for
--seconds, in three patterns: device->host alone, host->device alone, and both directions at once with the byte ratio of
the engine's step (actions up : results down).  All workers start together (a multiprocessing barrier) and the figure
per pattern is total bytes / the slowest worker's time.  Prints / writes one JSON object:
{"1": {"d2h_gbs": ..., "h2d_gbs": ..., "duplex_d2h_gbs": ..., "per_gpu": [...]}, "2": ...}.
The engine's e2e figure is quoted against `duplex_d2h_gbs` of the same GPU count (bench.py: e2e.link_gbs_measured).
"""
import argparse
import json
import multiprocessing as mp
import os
import sys
import time


def worker(dev, nbytes, seconds, up_ratio, barrier, q):
    import torch
    torch.cuda.set_device(dev)
    n_buf = 4
    host = [torch.empty(nbytes, dtype=torch.uint8).pin_memory() for _ in range(n_buf)]
    devb = [torch.empty(nbytes, dtype=torch.uint8, device=f"cuda:{dev}") for _ in range(n_buf)]
    up_bytes = max(256, int(nbytes * up_ratio) // 256 * 256)
    s_down, s_up = torch.cuda.Stream(dev), torch.cuda.Stream(dev)
    for h in host:
        h.fill_(1)
    torch.cuda.synchronize(dev)
    out = {}

    def run(down, up):
        for _ in range(2):  # warm-up
            if down:
                with torch.cuda.stream(s_down):
                    host[0].copy_(devb[0], non_blocking=True)
            if up:
                with torch.cuda.stream(s_up):
                    devb[1][:up_bytes].copy_(host[1][:up_bytes], non_blocking=True)
        torch.cuda.synchronize(dev)
        barrier.wait()
        t0 = time.perf_counter()
        copied_down = copied_up = 0
        k = 0
        while time.perf_counter() - t0 < seconds:
            for _ in range(4):  # keep a few copies queued, one cudaMemcpyAsync per buffer
                if down:
                    with torch.cuda.stream(s_down):
                        host[k % n_buf].copy_(devb[k % n_buf], non_blocking=True)
                    copied_down += nbytes
                if up:
                    with torch.cuda.stream(s_up):
                        devb[(k + 2) % n_buf][:up_bytes].copy_(host[(k + 2) % n_buf][:up_bytes], non_blocking=True)
                    copied_up += up_bytes
                k += 1
            s_down.synchronize()
            s_up.synchronize()
        torch.cuda.synchronize(dev)
        dt = time.perf_counter() - t0
        barrier.wait()
        return copied_down, copied_up, dt

    d, _, t = run(True, False)
    out["d2h"] = (d, t)
    up_bytes_full = up_bytes
    up_bytes = nbytes  # host->device alone: full-size buffers
    _, u, t = run(False, True)
    out["h2d"] = (u, t)
    up_bytes = up_bytes_full
    d, u, t = run(True, True)
    out["duplex"] = (d, u, t)
    q.put((dev, out))


def measure(g, nbytes, seconds, up_ratio):
    ctx = mp.get_context("spawn")
    barrier = ctx.Barrier(g)
    q = ctx.Queue()
    procs = [ctx.Process(target=worker, args=(d, nbytes, seconds, up_ratio, barrier, q)) for d in range(g)]
    for p in procs:
        p.start()
    res = dict(q.get() for _ in range(g))
    for p in procs:
        p.join()
    agg = {}
    tmax = max(res[d]["d2h"][1] for d in res)
    agg["d2h_gbs"] = sum(res[d]["d2h"][0] for d in res) / tmax / 1e9
    tmax = max(res[d]["h2d"][1] for d in res)
    agg["h2d_gbs"] = sum(res[d]["h2d"][0] for d in res) / tmax / 1e9
    tmax = max(res[d]["duplex"][2] for d in res)
    agg["duplex_d2h_gbs"] = sum(res[d]["duplex"][0] for d in res) / tmax / 1e9
    agg["duplex_h2d_gbs"] = sum(res[d]["duplex"][1] for d in res) / tmax / 1e9
    agg["per_gpu_d2h_gbs"] = [res[d]["d2h"][0] / res[d]["d2h"][1] / 1e9 for d in sorted(res)]
    return agg


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", default="1,2,4,8")
    ap.add_argument("--mb", type=float, default=34.0)
    ap.add_argument("--seconds", type=float, default=1.0)
    ap.add_argument("--up-ratio", type=float, default=2.0 / 34.0, help="host->device bytes per device->host byte in the duplex pattern")
    ap.add_argument("--out", default=None)
    args = ap.parse_args()
    import torch
    ndev = torch.cuda.device_count()
    if ndev == 0:
        raise SystemExit("host_link_peak: no CUDA device")
    result = {"buffer_mb": args.mb, "seconds": args.seconds, "host_cpus": len(os.sched_getaffinity(0)),
              "how": "pinned host buffers, one cudaMemcpyAsync per buffer (torch copy_ non_blocking), one process per GPU"}
    for g in [int(x) for x in args.gpus.split(",")]:
        if g > ndev:
            continue
        result[str(g)] = measure(g, int(args.mb * 1e6) // 256 * 256, args.seconds, args.up_ratio)
    s = json.dumps(result)
    print(s)
    if args.out:
        os.makedirs(os.path.dirname(os.path.abspath(args.out)), exist_ok=True)
        with open(args.out, "w") as f:
            f.write(s + "\n")


if __name__ == "__main__":
    sys.exit(main())
