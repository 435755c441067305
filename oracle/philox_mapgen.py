"""TEST INFRASTRUCTURE ONLY -- host mirror of the device's Philox map generator.

The reference draws maps from Python's global `random` stream (plantos_env.py:338-372),
which a GPU cannot reproduce; the CUDA reset therefore keeps the reference's
CONSTRUCTION (O//3 clusters with centre in [2, G-3]^2 and size 2 or 3, P distinct plants
on free cells each thirsty with probability p, rover on a free non-plant cell) but takes
its draws from Philox4x32-10 keyed by (seed, global env id, episode).  This module restates
that generator (rl_env_b200/csrc/plantos_generic.cuh: reset_env_warp) in pure Python so
tests can demand bit-identical maps from the device; distributional agreement with the
reference generator is tested separately.
"""
from __future__ import annotations

import math
from typing import Tuple

import numpy as np

M0, M1 = 0xD2511F53, 0xCD9E8D57
W0, W1 = 0x9E3779B9, 0xBB67AE85
MASK = 0xFFFFFFFF


def philox4x32_10(counter, key):
    c0, c1, c2, c3 = counter
    k0, k1 = key
    for _ in range(10):
        p0 = M0 * c0
        p1 = M1 * c2
        c0, c1, c2, c3 = ((p1 >> 32) ^ c1 ^ k0) & MASK, p1 & MASK, ((p0 >> 32) ^ c3 ^ k1) & MASK, p0 & MASK
        k0 = (k0 + W0) & MASK
        k1 = (k1 + W1) & MASK
    return c0, c1, c2, c3


def map_draw(seed: int, genv: int, episode: int, stream: int, j: int):
    counter = (j & MASK, episode & MASK, genv & MASK, (((genv >> 32) * 4) + stream) & MASK)
    return philox4x32_10(counter, (seed & MASK, (seed >> 32) & MASK))


def bounded(w: int, n: int) -> int:
    return (w * n) >> 32


def generate_map(seed: int, genv: int, episode: int, grid_size: int, num_plants: int,
                 num_obstacles: int, thirsty_plant_prob: float = 0.7) -> Tuple[np.ndarray, Tuple[int, int]]:
    """-> (cells u8 [G,G] with codes 0/1/2/3, rover (x, y))."""
    g = grid_size
    cells = np.zeros((g, g), dtype=np.uint8)
    for k in range(num_obstacles // 3):
        d = map_draw(seed, genv, episode, 0, k)
        cx = 2 + bounded(d[0], g - 4)
        cy = 2 + bounded(d[1], g - 4)
        size = 2 + (d[2] >> 31)
        for dx in range(size):
            for dy in range(size):
                ox, oy = cx + dx - 1, cy + dy - 1
                if 0 <= ox < g and 0 <= oy < g:
                    cells[ox, oy] = 1
    thresh = int(math.floor(float(np.float32(thirsty_plant_prob)) * 4294967296.0))
    thresh = min(max(thresh, 0), 1 << 32)
    j = placed = 0
    while placed < num_plants:
        d = map_draw(seed, genv, episode, 1, j)
        j += 1
        cell = bounded(d[0], g * g)
        cx, cy = divmod(cell, g)
        if cells[cx, cy] == 0:
            cells[cx, cy] = 3 if d[1] < thresh else 2
            placed += 1
    j = 0
    while True:
        d = map_draw(seed, genv, episode, 2, j)
        j += 1
        cell = bounded(d[0], g * g)
        cx, cy = divmod(cell, g)
        if cells[cx, cy] == 0:
            return cells, (cx, cy)
