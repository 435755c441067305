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


# ---------------------------------------------------------------- maze generator (map_source = "maze")
class _MazeRng:
    """Word j of Philox stream 3 = the j-th random decision (csrc/plantos_generic.cuh: MazeRng)."""

    def __init__(self, seed: int, genv: int, episode: int):
        self.seed, self.genv, self.episode, self.j, self.buf = seed, genv, episode, 0, None

    def next(self) -> int:
        if self.j & 3 == 0:
            self.buf = map_draw(self.seed, self.genv, self.episode, 3, self.j >> 2)
        w = self.buf[self.j & 3]
        self.j += 1
        return w

    def chance(self, thresh: int) -> bool:
        return self.next() < thresh


P30, P40, P20 = 1288490188, 1717986918, 858993459      # floor(p * 2^32) for p = 0.3, 0.4, 0.2


def _rect(cells, x0, x1, y0, y1, value):
    g = cells.shape[0]
    cells[max(x0, 0):min(x1, g), max(y0, 0):min(y1, g)] = value


def _room(rng, cells, mx, my):
    bx, by = mx * 6 + 1, my * 6 + 1
    _rect(cells, bx, bx + 5, by, by + 5, 0)
    if rng.chance(P30):
        _rect(cells, bx + 5, bx + 7, by + 2, by + 4, 0)
    if rng.chance(P30):
        _rect(cells, bx + 2, bx + 4, by + 5, by + 7, 0)
    if rng.chance(P40):
        c = bounded(rng.next(), 4)
        px, py = bx + (4 if c & 1 else 0), by + (4 if c & 2 else 0)
        _rect(cells, px, px + 1, py, py + 1, 1)


def maze_obstacles(seed: int, genv: int, episode: int, grid_size: int) -> np.ndarray:
    """The maze carving of maze_generate (csrc/plantos_generic.cuh): cells u8 [G, G], 1 obstacle / 0 free."""
    g, m = grid_size, (grid_size - 1) // 6
    cells = np.ones((g, g), dtype=np.uint8)
    if m < 1:
        return cells
    rng = _MazeRng(seed, genv, episode)
    cx, cy = bounded(rng.next(), m), bounded(rng.next(), m)
    visited = np.zeros((m, m), dtype=bool)
    stack = [(cx, cy)]
    visited[cx, cy] = True
    _room(rng, cells, cx, cy)
    while stack:
        cx, cy = stack[-1]
        cand = [k for k, (dx, dy) in enumerate([(0, 1), (0, -1), (1, 0), (-1, 0)])
                if 0 <= cx + dx < m and 0 <= cy + dy < m and not visited[cx + dx, cy + dy]]
        if not cand:
            stack.pop()
            continue
        k = cand[bounded(rng.next(), len(cand))]
        dx, dy = [(0, 1), (0, -1), (1, 0), (-1, 0)][k]
        nx, ny = cx + dx, cy + dy
        if dx == 0:
            _rect(cells, cx * 6 + 1, cx * 6 + 6, min(cy, ny) * 6 + 1, max(cy, ny) * 6 + 7, 0)
        else:
            _rect(cells, min(cx, nx) * 6 + 1, max(cx, nx) * 6 + 7, cy * 6 + 1, cy * 6 + 6, 0)
        if rng.chance(P20):
            mx, my = (cx + nx) // 2, (cy + ny) // 2
            d = 1 if bounded(rng.next(), 2) else -1
            if dx == 0:
                _rect(cells, mx * 6 + 2 + d * 2, mx * 6 + 4 + d * 2, my * 6 + 2, my * 6 + 4, 0)
            else:
                _rect(cells, mx * 6 + 2, mx * 6 + 4, my * 6 + 2 + d * 2, my * 6 + 4 + d * 2, 0)
        _room(rng, cells, nx, ny)
        visited[nx, ny] = True
        stack.append((nx, ny))
    return cells


def generate_maze_map(seed: int, genv: int, episode: int, grid_size: int, num_plants: int,
                      num_obstacles: int, thirsty_plant_prob: float = 0.7) -> Tuple[np.ndarray, Tuple[int, int]]:
    """map_source = "maze": maze obstacles (cluster generator as the fork's fallback when fewer than P + 1 cells
    are free), then plants and rover exactly as in generate_map."""
    g = grid_size
    cells = maze_obstacles(seed, genv, episode, g)
    if int((cells == 0).sum()) < num_plants + 1:
        return generate_map(seed, genv, episode, grid_size, num_plants, num_obstacles, thirsty_plant_prob)
    return _place_philox(seed, genv, episode, cells, num_plants, thirsty_plant_prob)


def _place_philox(seed, genv, episode, cells, num_plants, thirsty_plant_prob):
    g = cells.shape[0]
    thresh = int(math.floor(float(np.float32(thirsty_plant_prob)) * 4294967296.0))
    thresh = min(max(thresh, 0), 1 << 32)
    j = placed = 0
    while placed < num_plants:
        d = map_draw(seed, genv, episode, 1, j)
        j += 1
        cx, cy = divmod(bounded(d[0], g * g), g)
        if cells[cx, cy] == 0:
            cells[cx, cy] = 3 if d[1] < thresh else 2
            placed += 1
    j = 0
    while True:
        d = map_draw(seed, genv, episode, 2, j)
        j += 1
        cx, cy = divmod(bounded(d[0], g * g), g)
        if cells[cx, cy] == 0:
            return cells, (cx, cy)
