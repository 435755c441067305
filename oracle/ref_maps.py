"""TEST INFRASTRUCTURE ONLY -- restatement of the reference's two map generators for map injection
(`PlantOSVecEnv.push_maps`) and as the yardstick of the device generators' statistics.

The simulator's own resets draw maps on the device with Philox (csrc/plantos_generic.cuh: clusters and,
with map_source="maze", the fork's maze algorithm); this module produces maps exactly the way the
reference produces them -- with Python's global `random` module.  Two generators:

  * `original_map`: the cluster generator of plantos_env.py:338-372 (also `_generate_map_original` of
    the Gradio fork, gradio-app/plantos_env_new.py:360-406);
  * `maze_map`: the fork's `'maze'` algorithm (gradio-app/plantos_env_new.py:408-604):
    randomised depth-first search on a (G-1)//6 meta grid carving 5x5 rooms and 5-wide corridors, with
    random room extensions, corner cuts and path bulges.

Seeded with `random.seed(s)` they consume the same draws in the same order as the reference and return
the same map (checked in tests/test_oracle_vs_reference.py).  Cell codes: 0 empty, 1 obstacle,
2 hydrated plant, 3 thirsty plant; x = row = first index.
"""
from __future__ import annotations

import random
from typing import List, Sequence, Set, Tuple

import numpy as np

Cell = Tuple[int, int]


def _place(grid_size: int, obstacles: Set[Cell], num_plants: int, thirsty_plant_prob: float, rng):
    """Plants and rover (plantos_env.py:356-372): sample plants from the free cells, thirsty with
    probability p, rover on a remaining free cell.  The free set is built exactly like the
    reference builds it, because `list(set)` order decides which cells the draws select."""
    available = set((x, y) for x in range(grid_size) for y in range(grid_size)) - obstacles
    if len(available) < num_plants + 1:
        raise ValueError(f"Not enough available positions ({len(available)}) to place "
                         f"{num_plants} plants and 1 rover.")
    plants = {}
    plant_positions = rng.sample(list(available), num_plants)
    for pos in plant_positions:
        plants[pos] = rng.random() < thirsty_plant_prob
    available -= set(plant_positions)
    rover = rng.choice(list(available))
    cells = np.zeros((grid_size, grid_size), dtype=np.uint8)
    for (x, y) in obstacles:
        cells[x, y] = 1
    for (x, y), thirsty in plants.items():
        cells[x, y] = 3 if thirsty else 2
    return cells, (int(rover[0]), int(rover[1]))


def original_map(grid_size: int, num_plants: int, num_obstacles: int, thirsty_plant_prob: float = 0.7, rng=random):
    """plantos_env.py:338-372: num_obstacles // 3 clusters of 2x2 or 3x3 obstacles."""
    obstacles: Set[Cell] = set()
    for _ in range(num_obstacles // 3):
        cx = rng.randint(2, grid_size - 3)
        cy = rng.randint(2, grid_size - 3)
        size = rng.choice([2, 3])
        for dx in range(size):
            for dy in range(size):
                ox, oy = cx + dx - size // 2, cy + dy - size // 2
                if 0 <= ox < grid_size and 0 <= oy < grid_size:
                    obstacles.add((ox, oy))
    return _place(grid_size, obstacles, num_plants, thirsty_plant_prob, rng)


def maze_map(grid_size: int, num_plants: int, num_obstacles: int = 0, thirsty_plant_prob: float = 0.7, rng=random):
    """gradio-app/plantos_env_new.py:408-477 (+ helpers :479-580).  Falls back to `original_map`
    like the reference when the maze leaves no room for the plants and the rover (:463-467)."""
    g = grid_size
    obstacles: Set[Cell] = set((x, y) for x in range(g) for y in range(g))
    meta_w = meta_h = (g - 1) // 6

    def clear(px, py):
        if 0 <= px < g and 0 <= py < g:
            obstacles.discard((px, py))

    def carve_room(mx, my):                                   # :479-517
        bx, by = mx * 6 + 1, my * 6 + 1
        for i in range(5):
            for j in range(5):
                clear(bx + i, by + j)
        if rng.random() < 0.3:                                # extend right
            for i in range(2):
                for j in range(2, 4):
                    clear(bx + 5 + i, by + j)
        if rng.random() < 0.3:                                # extend down
            for i in range(2, 4):
                for j in range(2):
                    clear(bx + i, by + 5 + j)
        if rng.random() < 0.4:                                # cut one corner
            cx, cy = rng.choice([(0, 0), (4, 0), (0, 4), (4, 4)])
            px, py = bx + cx, by + cy
            if 0 <= px < g and 0 <= py < g:
                obstacles.add((px, py))

    def carve_straight(cx, cy, nx, ny, width=5):              # :540-560
        if cx == nx:
            for my in range(min(cy, ny), max(cy, ny) + 1):
                for i in range(width):
                    for j in range(6):
                        clear(cx * 6 + 1 + i, my * 6 + 1 + j)
        else:
            for mx in range(min(cx, nx), max(cx, nx) + 1):
                for i in range(6):
                    for j in range(width):
                        clear(mx * 6 + 1 + i, cy * 6 + 1 + j)

    def carve_path(cx, cy, nx, ny, dx, dy):                   # :519-538 (only cardinal moves occur)
        carve_straight(cx, cy, nx, ny)
        if rng.random() < 0.2:                                # bulge, :562-580
            mx, my = (cx + nx) // 2, (cy + ny) // 2
            direction = rng.choice([-1, 1])
            for i in range(2):
                for j in range(2):
                    if dx == 0:
                        clear(mx * 6 + 2 + direction * 2 + i, my * 6 + 2 + j)
                    else:
                        clear(mx * 6 + 2 + i, my * 6 + 2 + direction * 2 + j)

    visited = np.zeros((meta_w, meta_h), dtype=bool)
    sx, sy = rng.randint(0, meta_w - 1), rng.randint(0, meta_h - 1)
    stack: List[Cell] = [(sx, sy)]
    visited[sx, sy] = True
    carve_room(sx, sy)
    while stack:                                              # randomised DFS, :434-457
        cx, cy = stack[-1]
        neighbours = []
        for dx, dy in [(0, 1), (0, -1), (1, 0), (-1, 0)]:
            nx, ny = cx + dx, cy + dy
            if 0 <= nx < meta_w and 0 <= ny < meta_h and not visited[nx, ny]:
                neighbours.append((nx, ny, dx, dy))
        if neighbours:
            nx, ny, dx, dy = rng.choice(neighbours)
            carve_path(cx, cy, nx, ny, dx, dy)
            carve_room(nx, ny)
            visited[nx, ny] = True
            stack.append((nx, ny))
        else:
            stack.pop()
    if g * g - len(obstacles) < num_plants + 1:
        return original_map(grid_size, num_plants, num_obstacles, thirsty_plant_prob, rng)
    return _place(grid_size, obstacles, num_plants, thirsty_plant_prob, rng)


def make_maps(kind: str, num_envs: int, episodes: int, grid_size: int, num_plants: int, num_obstacles: int,
              thirsty_plant_prob: float = 0.7, rng=random):
    """`episodes` maps for each of `num_envs` envs in the layout `push_maps` takes:
    cells u8 [N, E, G, G], rover i16 [N, E, 2].  `kind` = "original" | "maze"."""
    gen = {"original": original_map, "maze": maze_map}[kind]
    cells = np.zeros((num_envs, episodes, grid_size, grid_size), dtype=np.uint8)
    rover = np.zeros((num_envs, episodes, 2), dtype=np.int16)
    for i in range(num_envs):
        for e in range(episodes):
            cells[i, e], rover[i, e] = gen(grid_size, num_plants, num_obstacles, thirsty_plant_prob, rng)
    return cells, rover
