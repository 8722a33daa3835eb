"""BatchedSpaceEnv — gym.vector-style front end of the CUDA engine.

Mirrors the reference's env interface for N independent envs (SpaceshipEnv.reset/step/seed/
vector_field, spaceship_env.py:59-100) with the same observation, action and reward layout,
env ids and kwargs.  All state lives in one caller-owned torch uint8 block in HBM that the
C library addresses directly (struct-of-arrays, see sg_layout); torch is used for memory,
streams and torch.distributed only — every computation is in libspacegym_b200.so.
"""
import ctypes as C

import numpy as np
import torch

from . import spaces
from ._lib import SgHostLayout, SgLayout, check, lib
from .registry import make_config


def shard_range(total_envs, rank, world_size):
    """Contiguous global env range of one rank: results are shard-count invariant."""
    base, rem = divmod(int(total_envs), int(world_size))
    start = rank * base + min(rank, rem)
    return start, base + (1 if rank < rem else 0)


def bind_host_to_gpu(device_index):
    """Pin the calling process to the CPUs NVML reports as local to GPU `device_index` (intersected with the CPUs
    the process may use), so that the pinned staging buffers of the host-buffer path are first-touched on, and
    copied by threads of, the GPU's own NUMA node.  One process per GPU: call it before the first env is made.
    Returns the CPU set in effect (unchanged if NVML or the intersection gives nothing)."""
    import os
    try:
        import pynvml
        pynvml.nvmlInit()
        h = pynvml.nvmlDeviceGetHandleByIndex(int(device_index))
        words = pynvml.nvmlDeviceGetCpuAffinity(h, (os.cpu_count() + 63) // 64)
        local = {64 * w + b for w, m in enumerate(words) for b in range(64) if (int(m) >> b) & 1}
        allowed = os.sched_getaffinity(0)
        if local & allowed:
            os.sched_setaffinity(0, local & allowed)
    except Exception:  # no NVML / not permitted: keep the inherited affinity
        pass
    return os.sched_getaffinity(0)


def allreduce_stats(vals):
    """Sum the 4-double episode statistics over torch.distributed ranks — the engine's only collective
    (SURVEY.md §8e).  NCCL needs the tensor on the GPU; gloo (CPU tests) reduces it in place."""
    import torch.distributed as dist
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
        return vals
    if dist.get_backend() == "nccl":
        dev = torch.device("cuda", torch.cuda.current_device())
        t = vals.to(dev)
        dist.all_reduce(t)
        return t.cpu()
    t = vals.clone()
    dist.all_reduce(t)
    return t


class BatchedSpaceEnv:
    def __init__(self, env_id, num_envs, device="cuda:0", seed=0, first_env_index=0, copy=True, host_obs_dtype="float32",
                 check_actions=True, **kwargs):
        """copy=True returns fresh tensors / arrays from step() (gym.vector's default); copy=False returns the
        engine's persistent output buffers, which the next step overwrites.

        host_obs_dtype: dtype of the observation the HOST-buffer path (`step(numpy)`, `step_async/step_wait`) brings
        back.  Default float32 — the dtype `observation_space` declares (spaceship_env.py:110-111) and what a policy
        consumes; it halves the bytes that cross the host link.  "float64" returns the engine's full-precision rows.
        The device path (`step(torch)`) is governed by `obs_dtype=` as before (default float64).

        check_actions: assert `action_space.contains(action)` on every step like the reference (spaceship_env.py:71);
        pass False on a hot loop (it reads the whole action batch)."""
        self._h = None
        self.copy = bool(copy)
        self.check_actions = bool(check_actions)
        self._out = {}
        self.env_id = env_id
        self.cfg = make_config(env_id, **kwargs)
        self.num_envs = int(num_envs)
        self.device = torch.device(device)
        if self.device.type != "cuda":
            raise RuntimeError("spacegym_b200 runs on CUDA devices only (there is no CPU path)")
        if not torch.cuda.is_available():
            raise RuntimeError("no CUDA device available: spacegym_b200 has no CPU fallback")
        self._L = lib()
        self.seed_value = int(seed)
        self.first_env_index = int(first_env_index)
        self.obs_dim = self._L.sg_obs_dim(C.byref(self.cfg))
        self.obs_dtype = torch.float32 if self.cfg.obs_dtype == 1 else torch.float64
        if host_obs_dtype not in ("float32", "float64"):
            raise ValueError("host_obs_dtype must be 'float32' or 'float64'")
        self.host_obs_dtype = host_obs_dtype
        lay = SgLayout()
        check(self._L.sg_state_layout(C.byref(self.cfg), self.num_envs, C.byref(lay)))
        self.layout = lay
        dev_index = self.device.index if self.device.index is not None else torch.cuda.current_device()
        self.device = torch.device("cuda", dev_index)
        with torch.cuda.device(self.device):
            self._state = torch.zeros(lay.total_bytes, dtype=torch.uint8, device=self.device)
            h = C.c_void_p()
            check(self._L.sg_create(C.byref(self.cfg), self.num_envs, dev_index, self.seed_value,
                                    self.first_env_index, self._state.data_ptr(), C.byref(h)))
        self._h = h
        self._host = None
        self._host_want_obs = True
        self._term = None
        self._flip = 0
        self._init_spaces()

    # ---- spaces (spaceship_env.py:102-111,183-214; kepler.py:158-170) -------------------------
    def _init_spaces(self):
        if self.cfg.discrete:
            self.single_action_space = spaces.Discrete(6)
        else:
            ones = np.ones(2, dtype=np.float32)
            self.single_action_space = spaces.Box(-ones, ones)
        if self.cfg.task == 0:
            hi = [1.0, 1.0, 1.0, 1.0, np.inf, np.inf, 1.0] + 2 * (self.cfg.n_planets + 1) * [2 * np.sqrt(2)]
            hi = np.array(hi, dtype=np.float32)
            self.single_observation_space = spaces.Box(-hi, hi)
        elif self.cfg.task == 1:
            hi = np.array([1.0, 1.0, 1.0, 1.0, np.inf, np.inf, 1.0, 2 * np.pi, 0.7, 2])
            self.single_observation_space = spaces.Box(-hi, hi, np.dtype(np.float64))
        else:  # DoNotCrash: with_lidar=False, with_goal=False (do_not_crash.py:30-31)
            hi = np.array([1.0, 1.0, 1.0, 1.0, np.inf, np.inf, 1.0], dtype=np.float32)
            self.single_observation_space = spaces.Box(-hi, hi)
        self.action_space, self.observation_space = self.single_action_space, self.single_observation_space

    # ---- helpers --------------------------------------------------------------------------------
    def _stream(self):
        return torch.cuda.current_stream(self.device).cuda_stream

    def _ptr(self, t):
        return None if t is None else C.c_void_p(t.data_ptr())

    def _check_dev(self, t, dtype, shape):
        if t.device != self.device or t.dtype != dtype or tuple(t.shape) != tuple(shape) or not t.is_contiguous():
            raise AssertionError(f"expected contiguous {dtype} tensor of shape {tuple(shape)} on {self.device}, "
                                 f"got {t.dtype} {tuple(t.shape)} on {t.device}")

    def _field(self, off, rows, dtype):
        n = self.num_envs
        nbytes = rows * n * torch.empty((), dtype=dtype).element_size()
        return self._state[off:off + nbytes].view(dtype).view(rows, n)

    # ---- gym.vector API ---------------------------------------------------------------------------
    def seed(self, seed=None):
        """env.seed (spaceship_env.py:92-94, goal.py:74-77): re-key the per-env random streams (keyed by
        (seed, global env index)) and restart their counters; takes effect with the next reset / step.
        `seed(s); reset()` replays what an env constructed with seed=s does.  Returns [seed] like the reference."""
        if seed is None:
            seed = int.from_bytes(__import__("os").urandom(8), "little") >> 1
        self.seed_value = int(seed)
        check(self._L.sg_reseed(self._h, self.seed_value), self._h)
        return [self.seed_value]

    def reset(self, mask=None):
        """Reset all envs (or those whose mask entry is non-zero); returns the observation of EVERY env (the
        current one for envs outside the mask)."""
        obs = torch.empty((self.num_envs, self.obs_dim), dtype=self.obs_dtype, device=self.device)
        if mask is not None:
            mask = mask.to(device=self.device, dtype=torch.uint8).contiguous()
        check(self._L.sg_reset(self._h, self._ptr(mask), self._ptr(obs), self._stream()), self._h)
        return obs

    def enable_terminal_observation(self, enable=True):
        """Under auto-reset `step` returns the reset observation of an env that finished (gym 0.21 SyncVectorEnv); the
        reference's own loop sees the episode's last observation because the caller resets (spaceship_env.py:68-78).
        After this call the device-path `step` / `step_many` put that row into `info["terminal_observation"]`
        ([N, D], the observation dtype; only rows of envs whose `done` is set in the step are new — the tensor is
        persistent and the other rows keep what they held)."""
        if enable:
            self._term = torch.zeros((self.num_envs, self.obs_dim), dtype=self.obs_dtype, device=self.device)
            check(self._L.sg_set_terminal_obs(self._h, self._ptr(self._term)), self._h)
        else:
            check(self._L.sg_set_terminal_obs(self._h, None), self._h)
            self._term = None
        return self._term

    def _assert_actions(self, a, is_torch):
        """action_space.contains for the whole batch (spaceship_env.py:69-72): Box(-1, 1, (2,)) or Discrete(6)."""
        if self.cfg.discrete:
            idx = a[..., 0] if a.ndim >= 2 and a.shape[-1] == 2 else a
            ok = ((idx >= 0) & (idx <= 5) & (idx == (idx.floor() if is_torch else np.floor(idx)))).all()
            assert bool(ok), "action outside Discrete(6)"
        else:
            assert bool(((a >= -1) & (a <= 1)).all()), "action outside Box(-1, 1)"

    def step(self, actions, check_actions=None):
        """obs [N, D], reward [N], done [N] (bool), info — one env.step for every env.

        torch CUDA actions run the device path; numpy actions run the host-buffer path
        (sg_step_host_block_async) and return numpy arrays."""
        if isinstance(actions, np.ndarray):
            return self._step_host(actions, check_actions)
        obs, rew, done, info = self.step_many(actions[None] if actions.dim() == 2 else actions.view(1, -1, 2),
                                              check_actions)
        out = {"TimeLimit.truncated": info["TimeLimit.truncated"][0]}
        if "terminal_observation" in info:
            out["terminal_observation"] = info["terminal_observation"]
        return obs, rew[0], done[0], out

    def step_many(self, actions=None, check_actions=None, K=None, want_obs=True):
        """K env steps per call (split kernels, or one fused launch with step_mode="fused"). actions: f32 [K, N, 2]
        (discrete: [..., 0] holds the index) or
        None with K given for synthetic random actions. Returns obs after the last step, reward [K, N],
        done [K, N], info."""
        n = self.num_envs
        if actions is not None:
            K = actions.shape[0]
            self._check_dev(actions, torch.float32, (K, n, 2))
            if self.check_actions if check_actions is None else check_actions:
                self._assert_actions(actions, True)
        K = int(K)
        # two output sets per K, used alternately: with copy=False the arrays of the previous call stay valid
        # for one more call (what a rollout loop needs: it reads step t's results while step t+1 runs)
        key = (K, 0 if self.copy else self._flip)
        self._flip = 1 - self._flip
        if key not in self._out:
            if len(self._out) > 8:
                self._out.clear()
            self._out[key] = dict(obs=torch.empty((n, self.obs_dim), dtype=self.obs_dtype, device=self.device),
                                  rew=torch.empty((K, n), dtype=torch.float64, device=self.device),
                                  done=torch.empty((K, n), dtype=torch.bool, device=self.device),
                                  trunc=torch.empty((K, n), dtype=torch.bool, device=self.device))
        o = self._out[key]
        obs = o["obs"] if want_obs else None
        # done / truncated are written as 0 / 1 bytes straight into torch.bool storage: no conversion pass
        check(self._L.sg_step(self._h, self._ptr(actions), K, self._ptr(obs), self._ptr(o["rew"]), self._ptr(o["done"]),
                              self._ptr(o["trunc"]), self._stream()), self._h)
        rew, done, trunc = o["rew"], o["done"], o["trunc"]
        if self.copy:
            obs = obs.clone() if obs is not None else None
            rew, done, trunc = rew.clone(), done.clone(), trunc.clone()
        info = {"TimeLimit.truncated": trunc}
        if self._term is not None and want_obs:
            info["terminal_observation"] = self._term.clone() if self.copy else self._term
        return obs, rew, done, info

    def rollout_random(self, K):
        """K env steps with the engine's synthetic uniform actions; nothing but statistics comes back."""
        check(self._L.sg_step(self._h, None, int(K), None, None, None, None, self._stream()), self._h)

    # ---- host-buffer path (numpy in, numpy out): gym.vector step_async / step_wait -------------------
    # One pinned result block per env object, laid out by the C library (sg_host_layout): observation | reward |
    # flags.  A step is one host->device copy (actions), the step kernels, one device->host copy (the block).
    def _host_buffers(self):
        if self._host is None:
            n, D = self.num_envs, self.obs_dim
            f32 = self.host_obs_dtype == "float32"
            lay = SgHostLayout()
            check(self._L.sg_host_block_layout(self._h, 1 if f32 else 0, 1, C.byref(lay)), self._h)
            block = torch.empty(lay.total_bytes, dtype=torch.uint8).pin_memory()
            act = torch.empty((n, 2), dtype=torch.float32).pin_memory()
            bn = block.numpy()
            odt = np.float32 if f32 else np.float64
            self._host = dict(block=block, act=act, layout=lay, f32=f32)
            self._host_np = dict(
                act=act.numpy(),
                obs=bn[lay.obs:lay.obs + n * D * np.dtype(odt).itemsize].view(odt).reshape(n, D),
                rew=bn[lay.reward:lay.reward + n * 8].view(np.float64),
                flags=bn[lay.flags:lay.flags + n],
                done=np.empty(n, dtype=np.bool_), trunc=np.empty(n, dtype=np.bool_))
        return self._host, self._host_np

    @property
    def action_buffer(self):
        """Pinned float32 [N, 2] action buffer: write actions here and call step_async(None) to skip a copy."""
        return self._host_buffers()[1]["act"]

    def step_async(self, actions=None, check_actions=None, want_obs=True):
        """Upload `actions` (numpy [N, 2] float32, or [N] ints for discrete envs; a pinned float32 CPU torch tensor
        is uploaded in place; None = already written to `action_buffer`), step, and start downloading the result block on the handle's own stream (ordered after
        the work already enqueued on the current torch stream).  want_obs=False leaves the observation on the
        device (9 bytes per env come back instead of 9 + 4 D)."""
        hb, hn = self._host_buffers()
        act_ptr = self._ptr(hb["act"])
        if isinstance(actions, torch.Tensor):
            # a pinned float32 [N, 2] CPU tensor is uploaded from where it is (no staging copy); the caller keeps it
            # unchanged until step_wait() returns
            if not (actions.device.type == "cpu" and actions.is_pinned() and actions.dtype == torch.float32
                    and actions.is_contiguous() and tuple(actions.shape) == (self.num_envs, 2)):
                raise ValueError("step_async(torch tensor): expected a pinned contiguous float32 CPU tensor of shape "
                                 f"{(self.num_envs, 2)} (CUDA tensors go through step())")
            if self.check_actions if check_actions is None else check_actions:
                self._assert_actions(actions, True)
            act_ptr = self._ptr(actions)
        elif actions is not None:
            a = np.asarray(actions)
            if self.check_actions if check_actions is None else check_actions:
                self._assert_actions(a, False)
            if self.cfg.discrete and a.ndim == 1:
                hn["act"][:, 0] = a
                hn["act"][:, 1] = 0
            else:
                assert a.shape == (self.num_envs, 2), f"expected actions of shape {(self.num_envs, 2)}, got {a.shape}"
                np.copyto(hn["act"], a, casting="same_kind")
        self._host_want_obs = bool(want_obs)
        check(self._L.sg_step_host_block_async(self._h, act_ptr, self._ptr(hb["block"]),
                                               1 if hb["f32"] else 0, 1 if want_obs else 0, self._stream()), self._h)

    def step_wait(self):
        """Block until the results of step_async are in host memory. The returned arrays are views of the env's
        pinned result block (overwritten by the next step) unless copy=True."""
        check(self._L.sg_host_wait(self._h), self._h)
        hn = self._host_np
        if self._host_want_obs:
            obs, rew, flags = hn["obs"], hn["rew"], hn["flags"]
        else:  # without the observation the block is reward | flags from offset 0
            n = self.num_envs
            bn = self._host["block"].numpy()
            lay = SgHostLayout()
            check(self._L.sg_host_block_layout(self._h, 1 if self._host["f32"] else 0, 0, C.byref(lay)), self._h)
            obs, rew, flags = None, bn[lay.reward:lay.reward + n * 8].view(np.float64), bn[lay.flags:lay.flags + n]
        done, trunc = hn["done"], hn["trunc"]
        np.not_equal(flags, 0, out=done)   # bit 0 = done, bit 1 = truncated (implies done)
        np.greater(flags, 1, out=trunc)
        if self.copy:
            obs = None if obs is None else obs.copy()
            rew, done, trunc = rew.copy(), done.copy(), trunc.copy()
        return obs, rew, done, {"TimeLimit.truncated": trunc}

    def _step_host(self, actions, check_actions=None):
        self.step_async(actions, check_actions)
        return self.step_wait()

    def vector_field(self, actions):
        """SpaceshipEnv.vector_field (spaceship_env.py:96-100) for every env: [N, 6]."""
        self._check_dev(actions, torch.float32, (self.num_envs, 2))
        out = torch.empty((self.num_envs, 6), dtype=torch.float64, device=self.device)
        check(self._L.sg_vector_field(self._h, self._ptr(actions), self._ptr(out), self._stream()), self._h)
        return out

    # ---- state access (parity tests, checkpoints) ------------------------------------------------
    def state_views(self):
        """Zero-copy [rows, N] views of the struct-of-arrays state block."""
        lay, c = self.layout, self.cfg
        v = {"ship": self._field(lay.ship, 6, torch.float64),
             "ep_return": self._field(lay.ep_return, 1, torch.float64),
             "elapsed": self._field(lay.elapsed, 1, torch.int32),
             "rng_ctr": self._field(lay.rng_ctr, 1, torch.int32),
             "status": self._field(lay.status, 1, torch.int32)}
        if c.task == 0:
            v["planets"] = self._field(lay.planets, 2 * c.n_planets, torch.float64)
            v["goal"] = self._field(lay.goal, 2, torch.float64)
            v["col_shift"] = self._field(lay.col_shift, c.cols, torch.float64)
            v["free_tiles"] = self._field(lay.free_tiles, 4, torch.int64)
            v["tile_word"] = self._field(lay.tile_word, 1, torch.int32)
        elif c.task == 1:
            v["orbit"] = self._field(lay.orbit, 2, torch.float64)
        return v

    def get_state(self):
        """Env-major copies: ship [N, 6], planets [N, P, 2], goal [N, 2], orbit [N, 2] = (ecc, angle), ..."""
        torch.cuda.current_stream(self.device).synchronize()
        out = {}
        for k, t in self.state_views().items():
            t = t.t().contiguous()
            if k == "planets":
                t = t.view(self.num_envs, self.cfg.n_planets, 2)
            elif t.shape[1] == 1:
                t = t[:, 0].contiguous()
            out[k] = t
        out["step_count"] = int(self._L.sg_step_count(self._h))
        return out

    def set_state(self, state):
        views = self.state_views()
        for k, val in state.items():
            if k == "step_count":
                check(self._L.sg_set_step_count(self._h, int(val)), self._h)
                continue
            dst = views[k]
            if dst.numel() == 0:  # a field this configuration does not have (e.g. column shifts without a tiling)
                continue
            src = torch.as_tensor(val, device=self.device).to(dst.dtype)
            src = src.reshape(self.num_envs, -1).t()
            dst.copy_(src)

    # ---- statistics ---------------------------------------------------------------------------------
    def stats(self, allreduce=False):
        """{'return_sum', 'length_sum', 'episodes', 'goals'} since creation / the last stats_reset().
        allreduce=True sums over torch.distributed ranks (the only collective of the engine)."""
        buf = (C.c_double * 4)()
        check(self._L.sg_stats(self._h, buf, self._stream()), self._h)
        vals = torch.tensor(list(buf), dtype=torch.float64)
        if allreduce:
            with torch.cuda.device(self.device):
                vals = allreduce_stats(vals)
        return dict(zip(("return_sum", "length_sum", "episodes", "goals"), vals.tolist()))

    def stats_reset(self):
        check(self._L.sg_stats_reset(self._h, self._stream()), self._h)

    @property
    def kernel_launches(self):
        return int(self._L.sg_kernel_launches(self._h))

    @property
    def step_count(self):
        return int(self._L.sg_step_count(self._h))

    def close(self):
        if self._h is not None:
            torch.cuda.synchronize(self.device)
            self._L.sg_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


def make(env_id, num_envs, **kwargs):
    """gym.make analogue for the batched engine."""
    return BatchedSpaceEnv(env_id, num_envs, **kwargs)
