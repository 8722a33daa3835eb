"""spacegym_b200 — B200-native batched stepping engine for the Space-Gym environments.

Product package: CUDA kernels + C ABI (csrc/, include/spacegym_b200.h) and the Python host
layer that mirrors the reference's gym interface.  Nothing here imports the CPU oracle.
"""
from .registry import ENV_IDS, make_config  # noqa: F401
from .vector_env import BatchedSpaceEnv, allreduce_stats, bind_host_to_gpu, make, shard_range  # noqa: F401
from .replay import Renderer, TrajectoryRecorder, replay, write_png  # noqa: F401  (off the hot path: debugging aid)
from .single_env import SpaceEnv, make_single  # noqa: F401  (the reference's single-env gym.Env call sites)

__version__ = "0.1.0"
