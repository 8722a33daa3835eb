#!/usr/bin/env python
"""bench.py — env-steps/s of the batched Space-Gym stepping engine (BASELINE.json metric).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl b200|reference]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N ... bench.py --gpus N ...

Workload (BASELINE.json configs[2], the config the north-star target is quoted on):
GoalContinuous3P-v0, 262144 envs per step and per GPU, auto-reset with hexagonal-tiling
placement, TimeLimit 500, uniform random float32 actions.  A "step" is one env.step() of one
batch of 262144 envs (one `sg_step` launch, K=1: actions read from HBM, observation, reward
and done written to HBM).  Sixteen independent env batches (each on its own stream) are stepped
round-robin, so that the working set (16 x 100 MB state+outputs) exceeds the 126 MB L2 between revisits
and the latency-bound continuation / tail kernels of one batch overlap the hot kernel of another.
CUDA_DEVICE_MAX_CONNECTIONS is raised to 32 (unless the caller set it): with the default of 8 hardware
queues, more than 8 streams alias onto the same queue and serialise (measured: 8 batches 4.94e9,
16 batches 4.31e9 at 8 connections, 5.09e9 at 32).

Prints ONE JSON line (rank 0).  `value` = device-timed (CUDA events, max over ranks) whole-job
env-steps/s with inputs resident in HBM; `e2e` = the same through BatchedSpaceEnv.step with
host (pinned) numpy buffers, copies inside the timed region; `roofline` = algorithmic HBM
bytes of the step kernel over its measured launch time against the measured copy bandwidth,
`roofline_fp64` the same for algorithmic FP64 flops against a DFMA micro-benchmark run here;
`cpu_baseline` = the C oracle port on this box's host cores.

`--impl reference` times the CPU implementation (oracle port, all host threads) on the same
workload and prints the same line with "impl": "reference".
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

os.environ.setdefault("CUDA_DEVICE_MAX_CONNECTIONS", "32")  # before the CUDA context exists; see the module doc

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

ENV_ID = "GoalContinuous3P-v0"
ENV_KW = {}
ENVS_PER_STEP = 262144
N_BATCHES = 16
ACTION_SETS = 4
# SURVEY.md §8d contract figures for GoalContinuous3P-v0 (per env step, K=1, f64 observation)
ALGO_BYTES_PER_STEP = 329.0
ALGO_FLOPS_PER_STEP = 1215.0
METRIC = "env-steps/sec (batched, device-timed)"
UNIT = "env-steps/s"


def workload_name(n):
    """The workload both arms are quoted on (BASELINE.json configs[2])."""
    return f"{ENV_ID} {n} envs/step/GPU auto-reset hexagonal-tiling TimeLimit500 random f32 actions, K=1"


def read_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        try:
            return float(json.load(open(p))["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
        except Exception:
            pass
    return 6650.0, "fallback (B200_PROFILING.md)"


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled DURING the timed region."""
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index, self.rows, self.proc = index, [], None

    def start(self):
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "50", "-i",
                 str(self.index)], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._read, daemon=True)
            self.thread.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([x.strip() for x in line.split(",")])

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in self.rows:
            try:
                sm.append(float(r[0]))
                mx.append(float(r[1]))
                for nm, v in zip(names, r[3:7]):
                    if v.lower().startswith("active"):
                        reasons.add(nm)
            except Exception:
                continue
        sm.sort()
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


def cpu_baseline(total_budget_s=12.0, n_envs=65536):
    """The oracle port (plain C, OpenMP over envs) on the host cores, bounded sample of the same workload."""
    from oracle import sg_oracle
    cores = len(os.sched_getaffinity(0))
    cfg = sg_oracle.config_from_id(ENV_ID)
    eng = sg_oracle.Engine(cfg, n_envs, seed=1, n_threads=cores)
    eng.reset()
    eng.step(None, K=2)
    t0 = time.perf_counter()
    eng.step(None, K=4)
    probe = (time.perf_counter() - t0) / 4
    steps = max(4, min(2000, int(total_budget_s / max(probe, 1e-6))))
    t0 = time.perf_counter()
    eng.step(None, K=steps)
    dt = time.perf_counter() - t0
    return {"value": n_envs * steps / dt, "unit": UNIT, "cores": cores, "kind": "port",
            "sample": f"{ENV_ID}, {n_envs} envs x {steps} steps, auto-reset, in-engine random actions, "
                      f"oracle/sg_oracle.c with {cores} OpenMP threads ({dt:.1f} s)"}


def run_reference(args):
    """Reference arm: CPU implementation of the path on the host cores (rank 0 only)."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    import numpy as np
    from oracle import sg_oracle
    cores = len(os.sched_getaffinity(0))
    cfg = sg_oracle.config_from_id(ENV_ID)
    n = ENVS_PER_STEP
    eng = sg_oracle.Engine(cfg, n, seed=1, n_threads=cores)
    eng.reset()
    rng = np.random.RandomState(0)
    acts = [rng.uniform(-1, 1, size=(1, n, 2)).astype(np.float32) for _ in range(ACTION_SETS)]
    t0 = time.perf_counter()
    eng.step(acts[0])
    probe = time.perf_counter() - t0
    # bounded sample: shrink the batch if K steps of 262144 envs would take more than ~2 minutes
    if probe * (args.steps + args.warmup) > 120:
        n = max(4096, int(n * 120 / (probe * (args.steps + args.warmup))) // 4096 * 4096)
        eng = sg_oracle.Engine(cfg, n, seed=1, n_threads=cores)
        eng.reset()
        acts = [a[:, :n].copy() for a in acts]
    for w in range(args.warmup):
        eng.step(acts[w % ACTION_SETS])
    t0 = time.perf_counter()
    for k in range(args.steps):
        eng.step(acts[k % ACTION_SETS])
    dt = time.perf_counter() - t0
    value = n * args.steps / dt
    sample = f"{ENV_ID}, {n} envs per step, {args.steps} steps, host numpy actions, oracle port, {cores} threads"
    print(json.dumps({
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": dt / args.steps * 1e3, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": workload_name(ENVS_PER_STEP),
                   "step_path": "CPU oracle port (oracle/sg_oracle.c), OpenMP over envs",
                   "envs_per_step_sampled": n},
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": cores, "kind": "port", "sample": sample},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }), flush=True)


def run_b200(args):
    import numpy as np
    import torch
    import torch.distributed as dist

    import spacegym_b200 as sg

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device (spacegym_b200 has no CPU path; use --impl reference)")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    cpus = sg.bind_host_to_gpu(local) if world > 1 else os.sched_getaffinity(0)  # NUMA-local pinned buffers per rank
    # stdout carries exactly one JSON line: anything a library prints at the C level (NCCL's version banner when
    # NCCL_DEBUG is set in the environment) is sent to stderr; the line itself goes to the saved descriptor
    sys.stdout.flush()
    json_fd = os.dup(1)
    os.dup2(2, 1)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=dev)

    n = args.envs
    total = n * world
    # envs sharded by contiguous global index range, no data-path collective
    envs = [sg.make(ENV_ID, n, device=dev, seed=1000 + 100000 * args.seed + b, first_env_index=rank * n, **ENV_KW) for b in range(N_BATCHES)]
    gen = torch.Generator(device=dev)
    gen.manual_seed(1234 + 100000 * args.seed + rank)
    actions = [[(torch.rand((1, n, 2), device=dev, generator=gen) * 2 - 1).float().contiguous()
                for _ in range(ACTION_SETS)] for _ in range(N_BATCHES)]
    D = envs[0].obs_dim
    outs = [dict(obs=torch.empty((n, D), dtype=torch.float64, device=dev),
                 rew=torch.empty((1, n), dtype=torch.float64, device=dev),
                 done=torch.empty((1, n), dtype=torch.uint8, device=dev),
                 trunc=torch.empty((1, n), dtype=torch.uint8, device=dev)) for _ in range(N_BATCHES)]
    for e in envs:
        e.reset()
        if args.prewarm > 0:
            e.rollout_random(args.prewarm)  # untimed: steady-state mix of episode ages
    L = envs[0]._L
    stream = torch.cuda.current_stream(dev)
    # the env batches are independent: each one is stepped on its own stream, forked from / joined into the
    # stream that carries the timing events
    batch_streams = [torch.cuda.Stream(device=dev) for _ in range(N_BATCHES)]

    def one_step(k):
        b = k % N_BATCHES
        e, o = envs[b], outs[b]
        a = actions[b][(k // N_BATCHES) % ACTION_SETS]
        rc = L.sg_step(e._h, a.data_ptr(), 1, o["obs"].data_ptr(), o["rew"].data_ptr(), o["done"].data_ptr(),
                       o["trunc"].data_ptr(), batch_streams[b].cuda_stream)
        if rc != 0:
            raise RuntimeError(L.sg_last_error(e._h))

    def fork():
        for s in batch_streams:
            s.wait_stream(stream)

    def join():
        for s in batch_streams:
            stream.wait_stream(s)

    def barrier():
        torch.cuda.synchronize(dev)
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize(dev)

    # the clock sampler starts BEFORE the warm-up: nvidia-smi's own start-up stalls the GPU for ~100 ms
    sampler = ClockSampler(local) if rank == 0 else None
    if sampler:
        sampler.start()
        t_wait = time.time()
        while not sampler.rows and time.time() - t_wait < 5.0:
            time.sleep(0.05)
    fork()
    for k in range(max(args.warmup, 3)):
        one_step(k)
    join()
    launches0 = sum(e.kernel_launches for e in envs)
    barrier()
    if sampler:
        sampler.rows.clear()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    ev0.record(stream)
    fork()
    for k in range(args.steps):
        one_step(k)
    join()
    ev1.record(stream)
    barrier()
    ms = ev0.elapsed_time(ev1)
    clocks = sampler.stop() if sampler else None
    launches = sum(e.kernel_launches for e in envs) - launches0
    t = torch.tensor([ms], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms = float(t.item())
    value = total * args.steps / (ms * 1e-3)

    # ---- one vector env alone: a single handle stepped back to back on one stream (what a user with ONE
    # BatchedSpaceEnv of this size gets; its state and outputs fit the L2, so this is the warm-L2 figure) ----
    single = None
    if rank == 0:
        e0, o0 = envs[0], outs[0]
        s_steps = max(50, min(args.steps, 400))

        def single_step(k):
            a = actions[0][k % ACTION_SETS]
            rc = L.sg_step(e0._h, a.data_ptr(), 1, o0["obs"].data_ptr(), o0["rew"].data_ptr(), o0["done"].data_ptr(),
                           o0["trunc"].data_ptr(), stream.cuda_stream)
            if rc != 0:
                raise RuntimeError(L.sg_last_error(e0._h))
        for k in range(10):
            single_step(k)
        torch.cuda.synchronize(dev)
        sv0, sv1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        sv0.record(stream)
        for k in range(s_steps):
            single_step(k)
        sv1.record(stream)
        torch.cuda.synchronize(dev)
        s_ms = sv0.elapsed_time(sv1)
        single = {"value": n * s_steps / (s_ms * 1e-3), "unit": UNIT, "envs": n, "batches": 1, "steps": s_steps,
                  "us_per_step": s_ms * 1e3 / s_steps,
                  "how": "one handle, sg_step back to back on one stream, CUDA events, warm L2"}
    barrier()

    # ---- end to end: public API with host buffers (pinned), copies inside the timed region ----
    # gym.vector step_async / step_wait over the same env batches: every step uploads its actions
    # from host memory and downloads observation, reward, done and truncated into host memory.
    e2e_steps = max(N_BATCHES, min(args.steps, 200) // N_BATCHES * N_BATCHES)
    host_acts = [np.random.RandomState(7 + rank + i).uniform(-1, 1, size=(n, 2)).astype(np.float32)
                 for i in range(ACTION_SETS)]
    for e in envs:
        e.copy = False  # results are read in place from the pinned buffers
    for k in range(2 * N_BATCHES):
        envs[k % N_BATCHES].step_async(host_acts[k % ACTION_SETS])
        if k >= N_BATCHES - 1:
            envs[(k + 1) % N_BATCHES].step_wait()
    for b in range(1, N_BATCHES):
        envs[b].step_wait()
    barrier()
    t0 = time.perf_counter()
    checksum = 0.0
    for k in range(e2e_steps):
        b = k % N_BATCHES
        if k >= N_BATCHES:  # results of this batch's previous step: consume before reusing its buffers
            obs, rew, done, info = envs[b].step_wait()
            checksum += float(rew[0]) + float(obs[-1, 0])
        envs[b].step_async(host_acts[(k // N_BATCHES) % ACTION_SETS])
    for b in range(N_BATCHES):
        obs, rew, done, info = envs[b].step_wait()
        checksum += float(rew[0]) + float(obs[-1, 0])
    barrier()
    e2e_s = time.perf_counter() - t0
    te = torch.tensor([e2e_s], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(te, op=dist.ReduceOp.MAX)
    e2e_value = total * e2e_steps / float(te.item())
    h2d = n * 2 * 4
    d2h = n * (D * 8 + 8 + 1 + 1)

    # ---- per-kernel durations, live (CUDA events on the launching stream; SG_TIMING handle) -------------
    kt = None
    if rank == 0:
        import ctypes
        os.environ["SG_TIMING"] = "1"
        tenv = sg.make(ENV_ID, n, device=dev, seed=4242, copy=False, **ENV_KW)
        del os.environ["SG_TIMING"]
        tenv.reset()
        tenv.rollout_random(max(args.prewarm, 64))
        ta = actions[0][0]
        flush = torch.empty(160 * 1024 * 1024, dtype=torch.uint8, device=dev)  # > L2: cold inputs for every step
        buf = (ctypes.c_double * 6)()
        for k in range(60):
            if k == 10:
                L.sg_debug_kernel_times(tenv._h, buf, 1)
            flush.fill_(k & 255)
            tenv.step_many(ta)
        L.sg_debug_kernel_times(tenv._h, buf, 1)
        kt = {"hot_us": buf[0] * 1e3, "cont_us": buf[1] * 1e3, "tail_us": buf[2] * 1e3,
              "cont_frac": buf[3] / n, "event_frac": buf[4] / n, "post_frac": buf[5] / n,
              "how": "50 steps of one 262144-env handle, L2 flushed before every step, CUDA events around each kernel"}
        tenv.close()
        del flush

    # ---- end to end with float32 observations (the dtype observation_space declares, spaceship_env.py:110) ----
    e2e32 = None
    if not args.no_f32:
        envs32 = [sg.make(ENV_ID, n, device=dev, seed=2000 + b, first_env_index=rank * n, copy=False,
                          obs_dtype="float32", **ENV_KW) for b in range(N_BATCHES)]
        for e in envs32:
            e.reset()
        for rep in range(2):
            barrier()
            t0 = time.perf_counter()
            for k in range(e2e_steps):
                b = k % N_BATCHES
                if k >= N_BATCHES:
                    obs, rew, done, info = envs32[b].step_wait()
                    checksum += float(rew[0]) + float(obs[-1, 0])
                envs32[b].step_async(host_acts[(k // N_BATCHES) % ACTION_SETS])
            for b in range(N_BATCHES):
                obs, rew, done, info = envs32[b].step_wait()
                checksum += float(rew[0])
            barrier()
            dt32 = time.perf_counter() - t0
        t32 = torch.tensor([dt32], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(t32, op=dist.ReduceOp.MAX)
        e2e32 = {"value": total * e2e_steps / float(t32.item()), "unit": UNIT, "h2d_bytes_per_step": h2d,
                 "d2h_bytes_per_step": n * (D * 4 + 8 + 1 + 1), "steps": e2e_steps}
        for e in envs32:
            e.close()

    stats = envs[0].stats(allreduce=world > 1)  # the engine's only collective: episode statistics
    # integrity of what was just timed (outside the timed region): sticky per-env fault bits (step too small, lost
    # Brent bracket, free-list overflow, non-finite state) and non-finite ship states over every batch of this rank
    torch.cuda.synchronize(dev)
    faults = torch.zeros(2, dtype=torch.int64, device=dev)
    for e in envs:
        v = e.state_views()
        faults[0] += (v["status"] != 0).sum()
        faults[1] += (~torch.isfinite(v["ship"])).any(dim=0).sum()
    if world > 1:
        dist.all_reduce(faults)
    faults = [int(x) for x in faults.tolist()]

    if rank == 0:
        hbm_peak, peak_src = read_peaks()
        launch_s = ms * 1e-3 / args.steps  # the timed region is exactly `steps` sg_step calls (3 kernels each)
        achieved_gbs = ALGO_BYTES_PER_STEP * n / launch_s / 1e9
        # dominant kernel = hot_kernel: it finishes (1 - cont - event) of the env steps of a launch by itself;
        # only those are counted as its algorithmic work, over its own live-measured duration
        hot_units = n * (1.0 - kt["cont_frac"] - kt["event_frac"])
        hot_s = kt["hot_us"] * 1e-6
        hot_gbs = ALGO_BYTES_PER_STEP * hot_units / hot_s / 1e9
        hot_tf = ALGO_FLOPS_PER_STEP * hot_units / hot_s / 1e12
        tf = C_double()
        rc = L.sg_measure_fp64_peak(local, 300.0, tf.ref)
        fp64_peak = tf.value if rc == 0 else None
        fp64_ach = ALGO_FLOPS_PER_STEP * n / launch_s / 1e12
        traffic = None
        tp = os.path.join(ROOT, "profiles", "traffic.json")
        if os.path.exists(tp):
            try:
                traffic = json.load(open(tp)).get("dram_bytes_per_launch")
            except Exception:
                traffic = None
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
            "warmup": max(args.warmup, 3), "ms_per_step": ms / args.steps, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {"workload": workload_name(n),
                       "step_path": "one sg_step call per step: hot / continuation / tail kernels",
                       "envs_per_gpu_per_step": n, "total_envs_per_step": total, "seed": args.seed,
                       "l2": f"inputs larger than L2: {N_BATCHES} env batches stepped round-robin "
                             f"({N_BATCHES * n * (ALGO_BYTES_PER_STEP + 60) / 1e6:.0f} MB working set)",
                       "parallelism": f"env-sharded x{world}, no data-path collective"},
            "roofline": {"bound": "hbm", "achieved": hot_gbs, "peak": hbm_peak, "unit": "GB/s",
                         "frac": hot_gbs / hbm_peak, "traffic": traffic, "peak_source": peak_src,
                         "kernel": "sg::hot_kernel<Goal,3P,velocity,f64>", "algo_bytes_per_env_step": ALGO_BYTES_PER_STEP,
                         "units_per_launch": hot_units, "launch_ms": kt["hot_us"] * 1e-3,
                         "note": "dominant kernel, duration from CUDA events on its stream with L2 flushed per step; "
                                 "units = env steps it completes alone"},
            "roofline_fp64": {"bound": "fp64", "achieved": hot_tf, "peak": fp64_peak, "unit": "TFLOP/s",
                              "frac": (hot_tf / fp64_peak) if fp64_peak else None,
                              "algo_flops_per_env_step": ALGO_FLOPS_PER_STEP, "kernel": "sg::hot_kernel",
                              "peak_source": "sg_measure_fp64_peak (DFMA chains, this run)"},
            "roofline_whole_step": {"hbm": {"achieved": achieved_gbs, "peak": hbm_peak, "unit": "GB/s",
                                            "frac": achieved_gbs / hbm_peak},
                                    "fp64": {"achieved": fp64_ach, "peak": fp64_peak, "unit": "TFLOP/s",
                                             "frac": (fp64_ach / fp64_peak) if fp64_peak else None},
                                    "note": "all kernels of a step (hot+continuation+tail), from `value`"},
            "kernel_times": kt,
            "single_handle": single,
            "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                    "steps": e2e_steps, "api": "BatchedSpaceEnv.step_async(numpy)/step_wait() -> sg_step_host_async/sg_host_wait, all env batches in flight"},
            "e2e_f32obs": e2e32,
            "gpu_launches": launches * world,
            "host_cpus_rank0": len(cpus),
            "clocks": clocks,
            "episode_stats": stats,
            "integrity": {"envs_with_fault_bits": faults[0], "envs_with_non_finite_state": faults[1],
                          "envs_checked": total * N_BATCHES, "when": "after the timed region"},
        }
        if world == 1 and not args.no_cpu_baseline:
            line["cpu_baseline"] = cpu_baseline()
        sys.stdout.flush()
        os.write(json_fd, (json.dumps(line) + "\n").encode())
    for e in envs:
        e.close()
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


class C_double:
    def __init__(self):
        import ctypes
        self._v = ctypes.c_double(0.0)
        self.ref = ctypes.byref(self._v)

    @property
    def value(self):
        return self._v.value


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=4000)
    ap.add_argument("--warmup", type=int, default=20)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--seed", type=int, default=0, help="offsets the env and action seeds of the timed batches (SURVEY 8d: seeds 0, 1, 2, median)")
    ap.add_argument("--envs", type=int, default=ENVS_PER_STEP, help="envs per step per GPU")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--prewarm", type=int, default=128, help="untimed env steps per batch that bring the envs to their steady-state mix of episode ages before warm-up")
    ap.add_argument("--batches", type=int, default=None, help="independent env batches in flight (default 16)")
    ap.add_argument("--no-f32", action="store_true", help="skip the float32-observation end-to-end leg")
    ap.add_argument("--env-id", default=None, help="benchmark another registered id (default: the BASELINE workload)")
    ap.add_argument("--kwargs", default="{}", help="JSON dict of reference-style kwargs for --env-id")
    args = ap.parse_args()
    global ENV_ID, ENV_KW, N_BATCHES
    if args.batches:
        N_BATCHES = args.batches
    if args.env_id:
        ENV_ID = args.env_id
    ENV_KW = json.loads(args.kwargs)
    if args.impl == "reference":
        run_reference(args)
    else:
        run_b200(args)


if __name__ == "__main__":
    main()
