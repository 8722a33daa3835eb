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
GATE_MS = 2.0
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


def ref_python_worker(env_id, seconds, seed, max_steps):
    """`bench.py --ref-worker`: one unmodified reference env (baseline/_ref/gym_space behind the gym shim of
    oracle/gym), random actions, TimeLimit auto-reset (spaceship_env.py:59-78).  Prints one JSON line."""
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    sys.path.insert(0, os.path.join(ROOT, "baseline", "_ref"))
    import numpy as np
    import gym
    import gym_space  # noqa: F401
    env = gym.make(env_id)
    env.seed(seed)
    env.reset()
    rng = np.random.RandomState(seed)
    acts = rng.uniform(-1, 1, size=(4096, 2)).astype(np.float32)
    for k in range(20):  # warm-up (imports, first solve_ivp)
        _, _, done, _ = env.step(acts[k])
        if done:
            env.reset()
    steps, episodes = 0, 0
    t0 = time.perf_counter()
    while steps < max_steps and time.perf_counter() - t0 < seconds:
        _, _, done, _ = env.step(acts[steps & 4095])
        steps += 1
        if done:
            env.reset()
            episodes += 1
    print(json.dumps({"steps": steps, "seconds": time.perf_counter() - t0, "episodes": episodes}), flush=True)


def cpu_baseline_python(seconds=3.0):
    """The UNMODIFIED reference gym loop on the host cores: one env per worker process over
    sched_getaffinity cores, random actions, TimeLimit(500) auto-reset, each worker wall-clocked after its own
    import / warm-up; plus BASELINE configs[0] (GoalContinuous2P-v0, one env, 1000 steps)."""
    if not os.path.isdir(os.path.join(ROOT, "baseline", "_ref", "gym_space")):
        return {"unavailable": "baseline/_ref/gym_space not staged (python tools/install_ref.py in the authoring container)"}
    cores = len(os.sched_getaffinity(0))
    env = dict(os.environ, OMP_NUM_THREADS="1", OPENBLAS_NUM_THREADS="1", MKL_NUM_THREADS="1", CUDA_VISIBLE_DEVICES="")

    def run(env_id, workers, secs, max_steps):
        procs = [subprocess.Popen([sys.executable, os.path.abspath(__file__), "--ref-worker", env_id, str(secs), str(100 + w),
                                   str(max_steps)], stdout=subprocess.PIPE, stderr=subprocess.PIPE, text=True, env=env)
                 for w in range(workers)]
        res = []
        for p in procs:
            try:
                out, err = p.communicate(timeout=secs + 120)
            except subprocess.TimeoutExpired:
                p.kill()
                return None, "reference worker timed out"
            if p.returncode != 0:
                return None, (err.strip().splitlines() or ["reference worker failed"])[-1]
            res.append(json.loads(out.strip().splitlines()[-1]))
        return res, None

    res, err = run(ENV_ID, cores, seconds, 1 << 30)
    if err:
        return {"unavailable": err}
    steps = sum(r["steps"] for r in res)
    value = sum(r["steps"] / r["seconds"] for r in res)
    out = {"value": value, "unit": UNIT, "cores": cores, "kind": "_ref", "per_core": value / cores,
           "sample": f"{ENV_ID}, unmodified gym_space (baseline/_ref) behind oracle/gym, one env per worker process x "
                     f"{cores}, random f32 actions, TimeLimit auto-reset, {steps} env steps in {seconds:.0f} s per worker "
                     f"({sum(r['episodes'] for r in res)} episodes)"}
    res, err = run("GoalContinuous2P-v0", 1, 60.0, 1000)
    if not err:
        out["config0"] = {"workload": "GoalContinuous2P-v0 single env, random actions, 1000 steps on CPU (reference Python step)",
                          "value": res[0]["steps"] / res[0]["seconds"], "unit": UNIT, "cores": 1, "steps": res[0]["steps"],
                          "seconds": res[0]["seconds"]}
    return out


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
    sm_hz = getattr(torch.cuda.get_device_properties(dev), "clock_rate", 1965000) * 1e3  # kHz -> Hz, for the gate
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
    # Device-side gate in front of ev0: the stream spins for ~GATE_MS while the host enqueues ev0, the forks and the
    # first ~100 steps behind it, so the device-timed region measures the GPU executing K steps and not the host's
    # enqueue latency (the fork alone is 16 event record/wait pairs).  A long run is device-bound with or without it
    # (17 us of host time against ~45 us of device time per step); a 20-step run is 1 ms of device time in total.
    torch.cuda._sleep(int(GATE_MS * 1e-3 * sm_hz))
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
    # gym.vector step_async / step_wait over the same env batches: every step uploads its actions from pinned host
    # memory (one H2D copy) and downloads ONE packed result block (observation | reward | done/truncated flags) into
    # pinned host memory.  The observation comes back in float32 (what observation_space declares,
    # spaceship_env.py:110-111; BatchedSpaceEnv's host-path default) — `e2e_f64obs` is the same with float64 rows.
    # Time-bounded: whole rounds over the batches until E2E_SECONDS have passed.
    host_acts = [torch.from_numpy(np.random.RandomState(7 + rank + i).uniform(-1, 1, size=(n, 2)).astype(np.float32)).pin_memory()
                 for i in range(ACTION_SETS)]
    blk = n * (D * 4 + 8 + 1)

    def e2e_leg(batch_envs, seconds):
        for e in batch_envs:
            e.copy = False  # results are read in place from the pinned block
        nb = len(batch_envs)
        for k in range(2 * nb):  # warm-up: allocates the pinned blocks, two steps per batch
            if k >= nb:
                batch_envs[k % nb].step_wait()
            batch_envs[k % nb].step_async(host_acts[k % ACTION_SETS], check_actions=False)
        for e in batch_envs:
            e.step_wait()
        barrier()
        t0 = time.perf_counter()
        checksum, rounds = 0.0, 1
        for b in range(nb):
            batch_envs[b].step_async(host_acts[b % ACTION_SETS], check_actions=False)
        while True:
            more = time.perf_counter() - t0 < seconds  # decided once per round over the batches
            for b in range(nb):
                # results of this batch's previous step: consumed before its buffers are reused
                obs, rew, done, info = batch_envs[b].step_wait()
                checksum += float(rew[0]) + float(obs[-1, 0]) + float(done[-1])
                if more:
                    batch_envs[b].step_async(host_acts[(rounds + b) % ACTION_SETS], check_actions=False)
            if not more:
                break
            rounds += 1
        steps_done = rounds * nb
        torch.cuda.synchronize(dev)
        dt = time.perf_counter() - t0
        tt = torch.tensor([dt, float(steps_done)], dtype=torch.float64, device=dev)
        if world > 1:  # whole job: env steps of all ranks over the slowest rank's time
            tmax = tt.clone()
            dist.all_reduce(tmax, op=dist.ReduceOp.MAX)
            dist.all_reduce(tt, op=dist.ReduceOp.SUM)
            return n * float(tt[1].item()) / float(tmax[0].item()), int(tt[1].item()) // world
        return n * steps_done / dt, steps_done

    e2e_value, e2e_steps = e2e_leg(envs, args.e2e_seconds)
    h2d = n * 2 * 4
    d2h = int(envs[0]._host["layout"].total_bytes)  # the block that is copied (256-byte aligned parts)

    # plain pinned-copy ceiling of the same link, measured here: device->host copies of one result block per
    # cudaMemcpyAsync, all ranks at once, nothing else running (tools/host_link_peak.py has the 1/2/4/8-GPU sweep)
    link_host = [torch.empty(blk, dtype=torch.uint8).pin_memory() for _ in range(2)]
    link_dev = [torch.empty(blk, dtype=torch.uint8, device=dev) for _ in range(2)]
    link_stream = torch.cuda.Stream(device=dev)
    with torch.cuda.stream(link_stream):
        for i in range(4):
            link_host[i % 2].copy_(link_dev[i % 2], non_blocking=True)
    barrier()
    t0 = time.perf_counter()
    copies = 0
    while time.perf_counter() - t0 < 0.4:
        with torch.cuda.stream(link_stream):
            for i in range(4):
                link_host[i % 2].copy_(link_dev[i % 2], non_blocking=True)
        link_stream.synchronize()
        copies += 4
    dt_link = time.perf_counter() - t0
    tl = torch.tensor([copies * blk / dt_link / 1e9], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(tl, op=dist.ReduceOp.SUM)
    link_gbs = float(tl.item())
    del link_host, link_dev

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

    # ---- end to end with float64 observation rows in the host block (full-precision rows; 2x the bytes) ----
    e2e64 = None
    if not args.no_f32:
        envs64 = [sg.make(ENV_ID, n, device=dev, seed=2000 + b, first_env_index=rank * n, copy=False,
                          host_obs_dtype="float64", **ENV_KW) for b in range(N_BATCHES)]
        for e in envs64:
            e.reset()
        v64, s64 = e2e_leg(envs64, min(args.e2e_seconds, 0.5))
        e2e64 = {"value": v64, "unit": UNIT, "h2d_bytes_per_step": h2d,
                 "d2h_bytes_per_step": n * (D * 8 + 8 + 1), "steps": s64}
        for e in envs64:
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
        traffic, pipe = None, None
        tp = os.path.join(ROOT, "profiles", "traffic.json")
        if os.path.exists(tp):
            try:
                tj = json.load(open(tp))
                traffic = tj.get("dram_bytes_per_launch")
                # whole-step FP64-pipe use in the overlapped state: executed FP64-pipe warp instructions of the three
                # kernels of one step (ncu counters, profiles/) over the step time measured HERE, against 0.5 warp
                # instructions per cycle on each of the SM sub-partitions at the SM clock sampled during the run
                ipp = tj.get("fp64_pipe", {}).get("fp64_pipe_warp_instructions_per_step")
                props = torch.cuda.get_device_properties(dev)
                mhz = (clocks or {}).get("sm_mhz") or 1965.0
                if ipp and n == ENVS_PER_STEP and ENV_ID == "GoalContinuous3P-v0":
                    peak = props.multi_processor_count * 4 * 0.5 * mhz * 1e6
                    pipe = {"fp64_pipe_warp_instructions_per_step": ipp, "achieved_per_s": ipp / launch_s, "peak_per_s": peak,
                            "frac": ipp / launch_s / peak, "sm_mhz": mhz,
                            "how": "instruction counts from ncu (profiles/traffic.json), step time from this run"}
            except Exception:
                traffic = None
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
            "warmup": max(args.warmup, 3), "ms_per_step": ms / args.steps, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {"workload": workload_name(n),
                       "step_path": "one sg_step call per step: hot / continuation / tail kernels",
                       "envs_per_gpu_per_step": n, "total_envs_per_step": total, "seed": args.seed,
                       "timing": f"CUDA events on the stream the batch streams fork from / join into; the K steps are "
                                 f"enqueued behind a {GATE_MS:.0f} ms device-side gate (host enqueue latency is outside the device-timed region)",
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
            "roofline_fp64_pipe_whole_step": pipe,
            "kernel_times": kt,
            "single_handle": single,
            "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                    "steps": e2e_steps, "obs_dtype": "float32",
                    "link_gbs_measured": link_gbs, "frac": e2e_value * (d2h / n) / 1e9 / link_gbs,
                    "link_how": f"pinned device->host cudaMemcpyAsync of {blk} B blocks, {world} rank(s) at once, 0.4 s, this run",
                    "api": "BatchedSpaceEnv.step_async(pinned actions)/step_wait() -> sg_step_host_block_async/sg_host_wait: "
                           "one H2D copy, hot/continuation/tail kernels, one D2H copy of the packed result block per step; "
                           "all env batches in flight"},
            "e2e_f64obs": e2e64,
            "gpu_launches": launches * world,
            "host_cpus_rank0": len(cpus),
            "clocks": clocks,
            "episode_stats": stats,
            "integrity": {"envs_with_fault_bits": faults[0], "envs_with_non_finite_state": faults[1],
                          "envs_checked": total * N_BATCHES, "when": "after the timed region"},
        }
        if world == 1 and not args.no_cpu_baseline:
            line["cpu_baseline"] = cpu_baseline()
            line["cpu_baseline_python"] = cpu_baseline_python()
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
    if len(sys.argv) >= 6 and sys.argv[1] == "--ref-worker":
        return ref_python_worker(sys.argv[2], float(sys.argv[3]), int(sys.argv[4]), int(sys.argv[5]))
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
    ap.add_argument("--no-f32", action="store_true", help="skip the float64-observation end-to-end leg")
    ap.add_argument("--e2e-seconds", type=float, default=0.6, help="wall-clock length of the end-to-end leg")
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
