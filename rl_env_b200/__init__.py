"""rl_env_b200 -- B200-native batched simulator for the GROW-R / PlantOS gridworld.

Only the env-step hot path of GammaKing2000/RL-Env (plantos_env.py:125-372 behind an
SB3 VecEnv) lives here: `csrc/` holds the sm_100a kernels and the C ABI of
include/plantos.h, `vec_env.py` the Python mirror of the reference's VecEnv surface.
"""
from .vec_env import (PRESETS, GraphRollout, LazyInfos, MonitorCSV, PendingStats, PlantOSVecEnv,  # noqa: F401
                      all_reduce_stats, shard_range)
from ._native import PlantOSError  # noqa: F401


def make_sharded(total_envs: int, rank: int, world_size: int, local_device: int = 0, **kwargs) -> PlantOSVecEnv:
    """The slice of a `total_envs` job that rank `rank` of `world_size` owns: contiguous
    global env ids, Philox counters keyed by the GLOBAL id so the maps do not depend on
    how many GPUs the job is spread over."""
    start, count = shard_range(total_envs, rank, world_size)
    return PlantOSVecEnv(count, device=f"cuda:{local_device}", env_id_base=start, **kwargs)
