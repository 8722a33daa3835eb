#!/usr/bin/env python
"""BASELINE.json configs[4]: GoalContinuous4P-v0, 1M..64M envs sharded over 1 / 2 / 4 / 8 GPUs.

    python tools/sweep_config5.py --out gpurun_out/sweep_g1.jsonl                                 # one GPU
    python -m torch.distributed.run --nnodes=1 --nproc-per-node G --master-addr 127.0.0.1 --master-port 29511 \
        tools/sweep_config5.py --out gpurun_out/sweep_gG.jsonl                                    # G GPUs

One vector env (ONE handle per GPU, which splits itself into sub-batches above 786432 envs) of `total / G` envs per rank,
envs sharded by contiguous global index, no data-path collective.  For every total: reset, 64 untimed steps, then
  * device-timed: K `sg_step` calls with device-resident actions, CUDA events, max over ranks -> env-steps/s (whole job);
  * end to end:   step_async(pinned actions) / step_wait() for ~0.4 s (one H2D + one packed D2H block per step);
  * integrity:    sticky fault bits and non-finite ship states after the run.
Totals: 2^20, 2^22, 2^24, 2^26 (strong scaling: the same total over more GPUs) plus G x 2^20 / 2^22 / 2^24 (weak scaling:
the same envs per GPU).  Rank 0 appends one JSON object per total to --out and prints it.
"""
import argparse
import json
import os
import sys
import time

os.environ.setdefault("CUDA_DEVICE_MAX_CONNECTIONS", "32")
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

ENV_ID = "GoalContinuous4P-v0"


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--out", default=None)
    ap.add_argument("--totals", default=None, help="comma-separated log2 of the total env counts (default: the sweep)")
    ap.add_argument("--e2e-seconds", type=float, default=0.4)
    ap.add_argument("--max-per-gpu-log2", type=int, default=26)
    args = ap.parse_args()
    import numpy as np
    import torch
    import torch.distributed as dist

    import spacegym_b200 as sg

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        sg.bind_host_to_gpu(local)
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=dev)
    if args.totals:
        logs = sorted({int(x) for x in args.totals.split(",")})
    else:
        g = world.bit_length() - 1
        logs = sorted({20, 22, 24, 26, 20 + g, 22 + g, 24 + g})
    logs = [lg for lg in logs if lg - (world.bit_length() - 1) <= args.max_per_gpu_log2 and (1 << lg) % world == 0]

    def barrier():
        torch.cuda.synchronize(dev)
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize(dev)

    def allmax(x):
        t = torch.tensor([x], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    def allsum(x):
        t = torch.tensor([x], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.SUM)
        return float(t.item())

    for lg in logs:
        total = 1 << lg
        n = total // world
        env = sg.make(ENV_ID, n, device=dev, seed=77, first_env_index=rank * n, copy=False)
        L = env._L
        env.reset()
        env.rollout_random(64)
        D = env.obs_dim
        gen = torch.Generator(device=dev)
        gen.manual_seed(99 + rank)
        acts = [(torch.rand((1, n, 2), device=dev, generator=gen) * 2 - 1).float().contiguous() for _ in range(2)]
        obs = torch.empty((n, D), dtype=torch.float64, device=dev)
        rew = torch.empty((1, n), dtype=torch.float64, device=dev)
        done = torch.empty((1, n), dtype=torch.uint8, device=dev)
        trunc = torch.empty((1, n), dtype=torch.uint8, device=dev)
        stream = torch.cuda.current_stream(dev)
        K = max(10, min(2000, int(1.5e9 // n)))

        def step(k):
            rc = L.sg_step(env._h, acts[k & 1].data_ptr(), 1, obs.data_ptr(), rew.data_ptr(), done.data_ptr(),
                           trunc.data_ptr(), stream.cuda_stream)
            if rc != 0:
                raise RuntimeError(L.sg_last_error(env._h))

        for k in range(5):
            step(k)
        barrier()
        ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        ev0.record(stream)
        for k in range(K):
            step(k)
        ev1.record(stream)
        barrier()
        ms = allmax(ev0.elapsed_time(ev1))
        value = total * K / (ms * 1e-3)

        host_acts = [torch.from_numpy(np.random.RandomState(5 + rank + i).uniform(-1, 1, size=(n, 2)).astype(np.float32)).pin_memory()
                     for i in range(2)]
        for k in range(2):
            env.step_async(host_acts[k], check_actions=False)
            env.step_wait()
        barrier()
        t0 = time.perf_counter()
        steps, chk = 0, 0.0
        while True:
            env.step_async(host_acts[steps & 1], check_actions=False)
            o, r, d, _ = env.step_wait()
            chk += float(r[0]) + float(o[-1, 0])
            steps += 1
            if time.perf_counter() - t0 >= args.e2e_seconds:
                break
        dt = allmax(time.perf_counter() - t0)
        e2e = n * allsum(steps) / dt
        d2h = int(env._host["layout"].total_bytes)

        v = env.state_views()
        faults = allsum(float((v["status"] != 0).sum().item()))
        nonfinite = allsum(float((~torch.isfinite(v["ship"])).any(dim=0).sum().item()))
        st = env.stats(allreduce=world > 1)
        row = {"env_id": ENV_ID, "total_envs": total, "n_gpus": world, "envs_per_gpu": n, "steps": K, "ms_per_step": ms / K, "value": value,
               "unit": "env-steps/s", "e2e": e2e, "e2e_steps_rank_mean": allsum(steps) / world,
               "h2d_bytes_per_step": n * 8, "d2h_bytes_per_step": d2h, "envs_with_fault_bits": int(faults),
               "envs_with_non_finite_state": int(nonfinite), "episodes": st.get("episodes"),
               "hbm_gb_in_use_rank0": (lambda fr, tot: (tot - fr) / 1e9)(*torch.cuda.mem_get_info(dev))}
        if rank == 0:
            s = json.dumps(row)
            print(s, flush=True)
            if args.out:
                with open(args.out, "a") as f:
                    f.write(s + "\n")
        env.close()
        del env, obs, rew, done, trunc, acts, host_acts, v
        torch.cuda.empty_cache()
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
