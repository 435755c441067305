"""Host-side constant tables for the PlantOS kernels.

Every entry is evaluated in Python double precision with the reference's own expression,
then narrowed once, so that the device never evaluates cos/sin or a division:

  lidar_off[i, r-1] = (int(r*cos(a)), int(r*sin(a))), a = (2*pi*i)/C   plantos_env.py:261-267
  dist_tab[r]       = float32(r / R)                                     plantos_env.py:288
  pos_tab[x]        = float32(x / G)                                     plantos_env.py:295-296
  visit_tab[k]      = float32(min(k, 10) / 10.0)                         plantos_env.py:308
  reward_tab[k]     = R_STEP + X_k  (and + R_COMPLETE_EXPLORATION)       plantos_env.py:164-181

The C library computes the same tables on its own (`plantos_compute_tables`); the VecEnv
uploads these Python-evaluated ones and `tests/test_abi_host.py` checks both agree.
"""
from __future__ import annotations

import math
from typing import Dict

import numpy as np

# order = PLANTOS_RW_* in include/plantos.h
REWARD_KEYS = ("r_exploration", "r_revisit", "r_invalid", "r_goal", "r_water_empty", "r_mistake")

# PlantOSEnv's active reward set (plantos_env.py:76-83)
DEFAULT_REWARDS: Dict[str, float] = dict(
    r_goal=20, r_mistake=-10, r_invalid=-5, r_water_empty=-5, r_step=-0.1,
    r_exploration=10, r_revisit=-1, r_complete_exploration=50)


def lidar_offsets(lidar_channels: int, lidar_range: int) -> np.ndarray:
    off = np.zeros((lidar_channels, lidar_range, 2), dtype=np.int8)
    for i in range(lidar_channels):
        angle = (2 * math.pi * i) / lidar_channels
        for r in range(1, lidar_range + 1):
            off[i, r - 1, 0] = int(r * math.cos(angle))
            off[i, r - 1, 1] = int(r * math.sin(angle))
    return off


def distance_table(lidar_range: int) -> np.ndarray:
    return np.array([r / lidar_range for r in range(lidar_range + 1)], dtype=np.float32)


def position_table(grid_size: int) -> np.ndarray:
    return np.array([x / grid_size for x in range(grid_size)], dtype=np.float32)


def visit_table() -> np.ndarray:
    return np.array([min(k, 10) / 10.0 for k in range(11)], dtype=np.float32)


def reward_table(rewards: Dict[str, float]) -> np.ndarray:
    """float64 [12]: entries 0..5 without, 6..11 with the completion bonus."""
    out = np.zeros(2 * len(REWARD_KEYS), dtype=np.float64)
    for k, key in enumerate(REWARD_KEYS):
        reward = rewards["r_step"]
        reward += rewards[key]
        out[k] = reward
        reward += rewards["r_complete_exploration"]
        out[len(REWARD_KEYS) + k] = reward
    return out
