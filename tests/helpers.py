"""Shared helpers for the GPU parity tests: golden/oracle <-> engine state conversion."""
import numpy as np
import torch

from golden_cases import CASES, TOL, close  # noqa: F401


def engine_kwargs(case):
    """Golden case -> (env id, kwargs) for spacegym_b200.make (reference-style kwargs)."""
    env_id, kw = CASES[case]
    kw = dict(kw)
    if "randomize" in kw:
        kw["randomize"] = bool(kw["randomize"])
    return env_id, kw


def oracle_cfg(oracle, case, **extra):
    env_id, kw = CASES[case]
    kw = dict(kw)
    kw.update(extra)
    return oracle.config_from_id(env_id, **kw)


def pack_tiles(envs):
    """oracle so_env tile fields -> engine tile_word / free_tiles words."""
    n = len(envs)
    goal_tile = np.where(envs["goal_tile"] < 0, 63, envs["goal_tile"]).astype(np.int64)
    word = (envs["case_b"].astype(np.int64) | (envs["flip_xy"].astype(np.int64) << 1)
            | (envs["ship_tile"].astype(np.int64) << 2) | (goal_tile << 8) | (envs["n_free"].astype(np.int64) << 14))
    fw = np.zeros((n, 4), dtype=np.uint64)
    free = envs["free_tiles"].astype(np.uint64)
    for k in range(64):
        live = (k < envs["n_free"]).astype(np.uint64)
        fw[:, k >> 4] |= (free[:, k] & np.uint64(15)) * live << np.uint64((k & 15) * 4)
    return word.astype(np.int32), fw.view(np.int64)


def oracle_state_to_engine(cfg, envs):
    """numpy structured so_env array -> dict accepted by BatchedSpaceEnv.set_state."""
    st = {"ship": envs["y"].copy(), "elapsed": envs["elapsed"].astype(np.int32),
          "rng_ctr": envs["rng_ctr"].astype(np.int64).astype(np.int32), "ep_return": envs["ep_return"].copy()}
    if cfg.task == 0:
        P = cfg.n_planets
        st["planets"] = envs["planets"][:, :2 * P].copy()
        st["goal"] = envs["goal"].copy()
        if cfg.cols > 0:  # the one-planet GoalEnv has no tiling (goal.py:79-82)
            st["col_shift"] = envs["col_shift"][:, :cfg.cols].copy()
            st["tile_word"], st["free_tiles"] = pack_tiles(envs)
    elif cfg.task == 1:
        st["orbit"] = np.stack([envs["orbit_ecc"], envs["orbit_angle"]], axis=1)
    return st


def compare_engine_to_oracle(cfg, dev_state, envs, tol=TOL):
    """Per-env boolean: engine state equals the oracle's so_env (floats to tol, integers exactly)."""
    ok = close(dev_state["ship"].cpu().numpy(), envs["y"], tol).all(axis=1)
    ok &= dev_state["elapsed"].cpu().numpy() == envs["elapsed"]
    ok &= dev_state["rng_ctr"].cpu().numpy().astype(np.uint32) == envs["rng_ctr"]
    ok &= close(dev_state["ep_return"].cpu().numpy(), envs["ep_return"], tol)
    if cfg.task == 0:
        P = cfg.n_planets
        n = len(envs)
        ok &= close(dev_state["planets"].cpu().numpy().reshape(n, 2 * P), envs["planets"][:, :2 * P], tol).all(axis=1)
        ok &= close(dev_state["goal"].cpu().numpy(), envs["goal"], tol).all(axis=1)
        if cfg.cols > 0:  # tile state exists only with a hexagonal tiling (n_planets >= 2)
            ok &= close(dev_state["col_shift"].cpu().numpy().reshape(n, -1), envs["col_shift"][:, :cfg.cols], tol).all(axis=1)
            word, fw = pack_tiles(envs)
            ok &= dev_state["tile_word"].cpu().numpy() == word
            ok &= (dev_state["free_tiles"].cpu().numpy().reshape(n, 4) == fw).all(axis=1)
    elif cfg.task == 1:
        ok &= close(dev_state["orbit"].cpu().numpy(),
                    np.stack([envs["orbit_ecc"], envs["orbit_angle"]], axis=1), tol).all(axis=1)
    return ok


def random_actions(rng, K, n, discrete):
    if discrete:
        a = np.zeros((K, n, 2), dtype=np.float32)
        a[..., 0] = rng.randint(0, 6, size=(K, n))
        return a
    return rng.uniform(-1, 1, size=(K, n, 2)).astype(np.float32)


def to_dev(a, device="cuda:0"):
    return torch.as_tensor(np.ascontiguousarray(a), device=device)
