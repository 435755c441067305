"""TEST INFRASTRUCTURE ONLY -- ctypes wrapper of oracle/plantos_oracle.c (the fast CPU checker)."""
from __future__ import annotations

import ctypes as C
import os
import subprocess
from typing import Optional

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
SRC = os.path.join(HERE, "plantos_oracle.c")
LIB = os.path.join(HERE, "libplantos_oracle.so")
SC_COUNT = 11


def build(force: bool = False) -> str:
    if force or not os.path.exists(LIB) or os.path.getmtime(LIB) < os.path.getmtime(SRC):
        subprocess.run(["gcc", "-O2", "-shared", "-fPIC", "-o", LIB, SRC, "-lm"], check=True)
    return LIB


_lib: Optional[C.CDLL] = None


def _load() -> C.CDLL:
    global _lib
    if _lib is None:
        lib = C.CDLL(build())
        vp = C.c_void_p
        lib.po_create.restype = vp
        lib.po_create.argtypes = [C.c_int] * 7
        lib.po_destroy.argtypes = [vp]
        lib.po_set_rewards.argtypes = [vp, vp]
        lib.po_set_maps.argtypes = [vp, vp, vp, C.c_int]
        lib.po_reset.argtypes = [vp, vp]
        lib.po_step.argtypes = [vp] * 10
        lib.po_reset_one.argtypes = [vp, C.c_int, vp, C.c_int, C.c_int, vp]
        lib.po_get_state.argtypes = [vp, vp, vp, vp]
        lib.po_lidar_offsets.argtypes = [C.c_int, C.c_int, vp]
        _lib = lib
    return _lib


def lidar_offsets(channels: int, rng: int) -> np.ndarray:
    out = np.zeros((channels, rng, 2), np.int8)
    _load().po_lidar_offsets(channels, rng, out.ctypes.data)
    return out


class COracle:
    """n envs stepped in index order with SB3 auto-reset; maps are injected."""

    def __init__(self, n: int, grid_size: int, num_plants: int, num_obstacles: int,
                 lidar_range: int, lidar_channels: int, max_steps: int = 1000):
        self.lib = _load()
        self.n, self.g, self.d = n, grid_size, 5 * lidar_channels + 27
        self.h = self.lib.po_create(n, grid_size, num_plants, num_obstacles, lidar_range, lidar_channels, max_steps)
        self._maps = None
        self.obs = np.zeros((n, self.d), np.float32)
        self.reward = np.zeros(n, np.float64)
        self.terminated = np.zeros(n, np.uint8)
        self.truncated = np.zeros(n, np.uint8)
        self.terminal_obs = np.zeros((n, self.d), np.float32)
        self.ep_return = np.zeros(n, np.float64)
        self.ep_len = np.zeros(n, np.int32)
        self.term_sc = np.zeros((SC_COUNT, n), np.int32)

    def __del__(self):
        if getattr(self, "h", None):
            self.lib.po_destroy(self.h)
            self.h = None

    def set_maps(self, cells: np.ndarray, rover: np.ndarray) -> None:
        cells = np.ascontiguousarray(cells, np.uint8)
        rover = np.ascontiguousarray(rover, np.int16)
        self._maps = (cells, rover)  # the C side borrows the buffers
        self.lib.po_set_maps(self.h, cells.ctypes.data, rover.ctypes.data, cells.shape[1])

    def reset(self) -> np.ndarray:
        self.lib.po_reset(self.h, self.obs.ctypes.data)
        return self.obs

    def reset_one(self, i: int, cells: np.ndarray, rover) -> None:
        cells = np.ascontiguousarray(cells, np.uint8)
        self.lib.po_reset_one(self.h, i, cells.ctypes.data, int(rover[0]), int(rover[1]),
                              self.obs[i].ctypes.data)

    def step(self, actions: np.ndarray):
        actions = np.ascontiguousarray(actions, np.int64)
        self.lib.po_step(self.h, actions.ctypes.data, self.obs.ctypes.data, self.reward.ctypes.data,
                         self.terminated.ctypes.data, self.truncated.ctypes.data,
                         self.terminal_obs.ctypes.data, self.ep_return.ctypes.data, self.ep_len.ctypes.data,
                         self.term_sc.ctypes.data)
        return self.obs, self.reward, self.terminated.astype(bool), self.truncated.astype(bool)

    def get_state(self):
        n, g = self.n, self.g
        cells = np.zeros((n, g, g), np.uint8)
        visits = np.zeros((n, g, g), np.int32)
        sc = np.zeros((SC_COUNT, n), np.int32)
        self.lib.po_get_state(self.h, cells.ctypes.data, visits.ctypes.data, sc.ctypes.data)
        return cells, visits, sc
