"""TEST INFRASTRUCTURE ONLY -- CPU restatement of the PlantOS env-step path.

This is the parity ORACLE for the CUDA path in `rl_env_b200/`.  It restates,
in plain Python + numpy, the algorithm of the reference's `PlantOSEnv`
(/root/reference/plantos_env.py:25-372) and of the SB3 `DummyVecEnv`+`Monitor`
loop the reference's trainers wrap it in (A2C_training.py:116-125,216-218).
Only `tests/`, `__graft_entry__.smoke()` and `bench.py`'s cpu_baseline /
`--impl reference` legs may import it; the product package never does.

Parity status: PINNED.  `tests/test_oracle_vs_reference.py` steps this port
and the unmodified reference (through `oracle/ref_shim.py`) side by side on
identical `random.seed`s and action streams and requires identical maps,
observations (bit-equal float32), rewards (bit-equal float64), flags and
info dicts; `tests/golden/*.npz` hold trajectories recorded from the
unmodified reference (generator: `tests/golden/make_golden.py`) and
`tests/test_oracle_golden.py` replays them through this port everywhere,
including on the GPU box where the reference checkout does not exist.

Hydrated-watering policy (plantos_env.py:213-222): the reference raises
TypeError when the rover waters an already-hydrated plant (the `return
self.R_MISTAKE` at :220 is unreachable).  This port returns the documented
R_MISTAKE (README.md:46; fixed fork gradio-app/plantos_env_new.py:236-245)
with no state change other than the step counter -- the same substitution
`ref_shim.ReferenceEnv` applies to the real reference.

Cell codes used for map exchange (one uint8 per cell, row-major [x][y], x is
the row / first index as in plantos_env.py:186-190):
    0 empty, 1 obstacle, 2 hydrated plant, 3 thirsty plant
which are the reference's LIDAR entity ids (plantos_env.py:20-23).
"""
from __future__ import annotations

import math
import random
import time
from typing import Any, Dict, List, Optional, Sequence, Tuple

import numpy as np

EMPTY, OBSTACLE, HYDRATED, THIRSTY = 0, 1, 2, 3

# name -> ctor kwargs; T = training preset (A2C_training.py:206-212),
# DFLT = ctor defaults (plantos_env.py:25-26), XL = stress preset.
PRESETS: Dict[str, Dict[str, int]] = {
    "T": dict(grid_size=25, num_plants=10, num_obstacles=12, lidar_range=6, lidar_channels=16),
    "DFLT": dict(grid_size=21, num_plants=8, num_obstacles=50, lidar_range=2, lidar_channels=10),
    "XL": dict(grid_size=64, num_plants=64, num_obstacles=600, lidar_range=32, lidar_channels=16),
}

INFO_KEYS = (
    "rover_position", "thirsty_plants", "hydrated_plants", "total_plants", "step_count",
    "explored_cells", "total_cells", "exploration_percentage", "lidar_range",
    "lidar_channels", "collided_with_wall", "total_collisions",
)


def lidar_offsets(channels: int, rng: int) -> np.ndarray:
    """int8 [C][R][2]: (dx, dy) of the r-th sample (r = 1..R) on ray i.

    plantos_env.py:261-267 -- angle = 2*pi*i/C; dx = int(r*cos), dy = int(r*sin)
    with Python's truncation toward zero; cos goes to x (row), sin to y (col).
    """
    out = np.zeros((channels, rng, 2), dtype=np.int8)
    for i in range(channels):
        angle = (2 * math.pi * i) / channels
        for r in range(1, rng + 1):
            out[i, r - 1, 0] = int(r * math.cos(angle))
            out[i, r - 1, 1] = int(r * math.sin(angle))
    return out


class PlantOSOracle:
    """One env.  Mirrors PlantOSEnv's public behaviour (plantos_env.py:25-372)."""

    def __init__(self, grid_size: int = 21, num_plants: int = 8, num_obstacles: int = 50,
                 lidar_range: int = 2, lidar_channels: int = 10,
                 thirsty_plant_prob: float = 0.7, max_steps: int = 1000,
                 literal_trig: bool = True):
        # plantos_env.py:31-36
        self.grid_size = grid_size
        self.num_plants = num_plants
        self.num_obstacles = num_obstacles
        self.lidar_range = lidar_range
        self.lidar_channels = lidar_channels
        self.thirsty_plant_prob = thirsty_plant_prob
        # plantos_env.py:45-57: 5 per ray + 2 position + 25 visit window
        self.obs_dim = lidar_channels * 5 + 2 + 25
        # plantos_env.py:76-83 (the active "DQN" reward set)
        self.R_GOAL = 20
        self.R_MISTAKE = -10
        self.R_INVALID = -5
        self.R_WATER_EMPTY = -5
        self.R_STEP = -0.1
        self.R_EXPLORATION = 10
        self.R_REVISIT = -1
        self.R_COMPLETE_EXPLORATION = 50
        self.max_steps = max_steps  # plantos_env.py:120
        # literal_trig=True evaluates cos/sin per sample like the reference does
        # (keeps the CPU-baseline cost profile honest); False uses the table.
        self.literal_trig = literal_trig
        self._offsets = lidar_offsets(lidar_channels, lidar_range).tolist()

        self.rover_pos: Optional[Tuple[int, int]] = None
        self.plants: Dict[Tuple[int, int], bool] = {}
        self.obstacles: set = set()
        self.explored_map: Optional[np.ndarray] = None
        self.visit_counts: Optional[np.ndarray] = None
        self.step_count = 0
        self.collided_with_wall = False
        self.completion_bonus_given = False
        self.total_collisions = 0
        self.mistake_steps = 0  # how often the hydrated-watering policy fired

    # ------------------------------------------------------------------ maps
    def generate_map(self) -> None:
        """Procedural map from the GLOBAL `random` stream (plantos_env.py:338-372).

        Draw order: per cluster randint, randint, choice; then sample(P);
        one random() per plant in sample order; one choice for the rover.  The
        set expressions are kept in the same shape so `list(set)` enumerates
        in the same order as in the reference for a given CPython.
        """
        g = self.grid_size
        self.obstacles = set()
        self.plants = {}
        for _ in range(self.num_obstacles // 3):
            cx = random.randint(2, g - 3)
            cy = random.randint(2, g - 3)
            size = random.choice([2, 3])
            for dx in range(size):
                for dy in range(size):
                    ox = cx + dx - size // 2
                    oy = cy + dy - size // 2
                    if 0 <= ox < g and 0 <= oy < g:
                        self.obstacles.add((ox, oy))
        free = set((x, y) for x in range(g) for y in range(g)) - self.obstacles
        if len(free) < self.num_plants + 1:
            raise ValueError(
                f"Not enough available positions ({len(free)}) to place "
                f"{self.num_plants} plants and 1 rover.")
        chosen = random.sample(list(free), self.num_plants)
        for pos in chosen:
            self.plants[pos] = random.random() < self.thirsty_plant_prob
        free -= set(chosen)
        self.rover_pos = random.choice(list(free))

    def inject_map(self, cells: np.ndarray, rover: Sequence[int]) -> None:
        """Install a recorded map: `cells` uint8 [G][G] of cell codes, rover (x, y)."""
        g = self.grid_size
        cells = np.asarray(cells).reshape(g, g)
        self.obstacles = set()
        self.plants = {}
        for x in range(g):
            for y in range(g):
                c = int(cells[x, y])
                if c == OBSTACLE:
                    self.obstacles.add((x, y))
                elif c == HYDRATED:
                    self.plants[(x, y)] = False
                elif c == THIRSTY:
                    self.plants[(x, y)] = True
        self.rover_pos = (int(rover[0]), int(rover[1]))

    def cell_plane(self) -> np.ndarray:
        g = self.grid_size
        plane = np.zeros((g, g), dtype=np.uint8)
        for (x, y) in self.obstacles:
            plane[x, y] = OBSTACLE
        for (x, y), thirsty in self.plants.items():
            plane[x, y] = THIRSTY if thirsty else HYDRATED
        return plane

    # ----------------------------------------------------------------- reset
    def reset(self, map_cells: Optional[np.ndarray] = None,
              rover: Optional[Sequence[int]] = None):
        """plantos_env.py:125-158.  With `map_cells` the map is injected instead
        of drawn from `random` (the mode the GPU path is compared in)."""
        self.step_count = 0
        self.collided_with_wall = False
        self.completion_bonus_given = False
        self.total_collisions = 0
        if map_cells is None:
            self.generate_map()
        else:
            self.inject_map(map_cells, rover)
        g = self.grid_size
        self.explored_map = np.zeros((g, g), dtype=np.int8)
        self.explored_map[self.rover_pos[0], self.rover_pos[1]] = 2
        self.visit_counts = np.zeros((g, g), dtype=np.int32)
        self.visit_counts[self.rover_pos[0], self.rover_pos[1]] = 1
        return self.observe(), self.info()

    # ------------------------------------------------------------------ step
    def step(self, action: int):
        """plantos_env.py:160-183."""
        self.step_count += 1
        reward = self.R_STEP
        if action < 4:
            reward += self._move(action)
        else:
            reward += self._water()
        obs = self.observe()
        info = self.info()
        terminated = bool(info["exploration_percentage"] >= 100)
        truncated = self.step_count >= self.max_steps
        if info["exploration_percentage"] >= 100 and not self.completion_bonus_given:
            reward += self.R_COMPLETE_EXPLORATION
            self.completion_bonus_given = True
        return obs, reward, terminated, truncated, info

    def _move(self, action: int) -> float:
        """plantos_env.py:185-211.  N, E, S, W on (x, y); plants are walkable."""
        dx, dy = ((-1, 0), (0, 1), (1, 0), (0, -1))[action]
        nx = self.rover_pos[0] + dx
        ny = self.rover_pos[1] + dy
        g = self.grid_size
        if 0 <= nx < g and 0 <= ny < g and (nx, ny) not in self.obstacles:
            fresh = self.visit_counts[nx, ny] == 0
            self.explored_map[self.rover_pos[0], self.rover_pos[1]] = 1
            self.rover_pos = (nx, ny)
            self.explored_map[nx, ny] = 2
            self.visit_counts[nx, ny] += 1
            return self.R_EXPLORATION if fresh else self.R_REVISIT
        self.collided_with_wall = True
        self.total_collisions += 1
        return self.R_INVALID

    def _water(self) -> float:
        """plantos_env.py:213-222 with the documented hydrated-plant result."""
        if self.rover_pos in self.plants:
            if self.plants[self.rover_pos]:
                self.plants[self.rover_pos] = False
                return self.R_GOAL
            self.mistake_steps += 1
            return self.R_MISTAKE
        return self.R_WATER_EMPTY

    # ----------------------------------------------------------- observation
    def lidar_hits(self) -> List[Tuple[int, int]]:
        """Per ray (distance r in 1..R, entity code) -- plantos_env.py:260-284."""
        g = self.grid_size
        rx, ry = self.rover_pos
        hits = []
        for i in range(self.lidar_channels):
            dist, kind = self.lidar_range, EMPTY
            if self.literal_trig:
                angle = (2 * math.pi * i) / self.lidar_channels
            for r in range(1, self.lidar_range + 1):
                if self.literal_trig:
                    cx = rx + int(r * math.cos(angle))
                    cy = ry + int(r * math.sin(angle))
                else:
                    off = self._offsets[i][r - 1]
                    cx = rx + off[0]
                    cy = ry + off[1]
                if not (0 <= cx < g and 0 <= cy < g):
                    dist, kind = r, OBSTACLE
                    break
                if (cx, cy) in self.obstacles:
                    dist, kind = r, OBSTACLE
                    break
                if (cx, cy) in self.plants:
                    dist, kind = r, (THIRSTY if self.plants[(cx, cy)] else HYDRATED)
                    break
            hits.append((dist, kind))
        return hits

    def observe(self) -> np.ndarray:
        """float32 [5C+27] -- plantos_env.py:251-315."""
        c = self.lidar_channels
        obs = np.zeros(self.obs_dim, dtype=np.float32)
        for i, (dist, kind) in enumerate(self.lidar_hits()):
            obs[5 * i] = dist / self.lidar_range
            obs[5 * i + 1 + kind] = 1.0
        rx, ry = self.rover_pos
        g = self.grid_size
        obs[5 * c] = rx / g
        obs[5 * c + 1] = ry / g
        base = 5 * c + 2
        for lx in range(5):
            for ly in range(5):
                gx = rx + lx - 2
                gy = ry + ly - 2
                if 0 <= gx < g and 0 <= gy < g:
                    obs[base + lx * 5 + ly] = min(self.visit_counts[gx, gy], 10) / 10.0
                else:
                    obs[base + lx * 5 + ly] = 1.0
        return obs

    def info(self) -> Dict[str, Any]:
        """plantos_env.py:317-336."""
        thirsty = sum(self.plants.values())
        explored = np.sum(self.explored_map > 0)
        total = self.grid_size * self.grid_size - len(self.obstacles)
        return {
            "rover_position": self.rover_pos,
            "thirsty_plants": thirsty,
            "hydrated_plants": len(self.plants) - thirsty,
            "total_plants": len(self.plants),
            "step_count": self.step_count,
            "explored_cells": explored,
            "total_cells": total,
            "exploration_percentage": (explored / total) * 100,
            "lidar_range": self.lidar_range,
            "lidar_channels": self.lidar_channels,
            "collided_with_wall": self.collided_with_wall,
            "total_collisions": self.total_collisions,
        }

    # ------------------------------------------------------------ state dump
    def export_state(self) -> Dict[str, Any]:
        """Integer state in the layout `plantos_get_state` returns (include/plantos.h)."""
        return {
            "cells": self.cell_plane(),
            "visits": self.visit_counts.astype(np.int32).copy(),
            "x": self.rover_pos[0], "y": self.rover_pos[1],
            "step_count": self.step_count,
            "explored_cells": int(np.sum(self.explored_map > 0)),
            "total_cells": self.grid_size ** 2 - len(self.obstacles),
            "thirsty_plants": int(sum(self.plants.values())),
            "total_collisions": self.total_collisions,
            "collided_with_wall": int(self.collided_with_wall),
            "completion_bonus_given": int(self.completion_bonus_given),
        }


def rollout_policy(env: "PlantOSOracle", u0: float, u1: float) -> int:
    """The MCTS planner's rollout policy restated (mcts_custom_trainer.py:168-216) with the two random
    draws supplied: `u0 < 0.7` selects `_exploration_heuristic` (least visited valid neighbour, first
    minimum in N, E, S, W order, :196-216), otherwise -- and when every move is blocked -- the action is
    `int(5 * u1)` (the reference draws np.random.randint(5))."""
    rnd = min(4, max(0, int(np.float32(u1) * np.float32(5.0))))
    if not (np.float32(u0) < np.float32(0.7)):
        return rnd
    x, y = env.rover_pos
    best, min_visits = None, None
    for action, (dx, dy) in enumerate([(-1, 0), (0, 1), (1, 0), (0, -1)]):
        nx, ny = x + dx, y + dy
        if 0 <= nx < env.grid_size and 0 <= ny < env.grid_size and (nx, ny) not in env.obstacles:
            visits = int(env.visit_counts[nx, ny])
            if min_visits is None or visits < min_visits:
                min_visits, best = visits, action
    return best if best is not None else rnd


class CurriculumOracle:
    """`CurriculumWrapper` restated -- both variants the reference ships:

      "a2c": A2C_training.py:37-109  (ctor call :121: thresholds 40 -> 100 by 10, 3 episodes per
             maze, reaching the threshold TERMINATES the episode, :98-100)
      "dqn": trainingCode.py:24-98   (ctor call :107: 30 -> 100 by 5, 50 episodes per maze, reaching
             the threshold only marks the maze completed, :88-89)

    What it really does (the map generator ignores `seed`, plantos_env.py:127 vs :344-372, so the
    "same maze" is a new map every time): `visit_counts` is carried over to the next episode unless
    the maze was completed or has been played `max_episodes_per_maze` times -- restored AFTER the
    reset observation was built, so that observation shows fresh counts and the rover's start cell
    is not counted -- while `explored_map` restarts every episode.  Rewards therefore use the
    persistent counts (`was_new`, plantos_env.py:197) and the exploration percentage does not."""

    VARIANTS = {
        "a2c": dict(initial_threshold=40.0, max_threshold=100.0, threshold_increment=10.0,
                    max_episodes_per_maze=3, terminate_on_threshold=True),
        "dqn": dict(initial_threshold=30.0, max_threshold=100.0, threshold_increment=5.0,
                    max_episodes_per_maze=50, terminate_on_threshold=False),
    }

    def __init__(self, env: "PlantOSOracle", variant: str = "a2c", **override):
        cfg = dict(self.VARIANTS[variant]); cfg.update(override)
        self.env = env
        self.maze_completed = False
        self.episodes_on_current_maze = 0
        self.persistent_visit_counts = None
        self.exploration_threshold = float(cfg["initial_threshold"])
        self.max_threshold = float(cfg["max_threshold"])
        self.threshold_increment = float(cfg["threshold_increment"])
        self.max_episodes_per_maze = int(cfg["max_episodes_per_maze"])
        self.terminate_on_threshold = bool(cfg["terminate_on_threshold"])

    def __getattr__(self, name):
        return getattr(self.env, name)

    def reset(self, map_cells=None, rover=None):
        self.episodes_on_current_maze += 1
        timeout = self.episodes_on_current_maze >= self.max_episodes_per_maze
        if self.maze_completed or timeout:
            if self.maze_completed:
                self.exploration_threshold = min(self.exploration_threshold + self.threshold_increment,
                                                 self.max_threshold)
            self.maze_completed = False
            self.episodes_on_current_maze = 0
            obs, info = self.env.reset(map_cells, rover)
            self.persistent_visit_counts = None
        else:
            obs, info = self.env.reset(map_cells, rover)
            if self.persistent_visit_counts is not None:
                self.env.visit_counts = self.persistent_visit_counts.copy()
            else:
                self.persistent_visit_counts = self.env.visit_counts.copy()
        return obs, info

    def step(self, action: int):
        obs, reward, terminated, truncated, info = self.env.step(action)
        if info["exploration_percentage"] >= self.exploration_threshold:
            self.maze_completed = True
            if self.terminate_on_threshold:
                terminated = True
        if self.persistent_visit_counts is not None:
            self.persistent_visit_counts = self.env.visit_counts.copy()
        return obs, reward, terminated, truncated, info


class OracleVecEnv:
    """`DummyVecEnv([Monitor(PlantOSEnv(**kw))] * n)` restated.

    SB3 is third-party (stable-baselines3==2.2.1, requirements.txt:6) and not
    vendored; the reference holds no test on this boundary, so this class pins
    the behaviour by spelling it out: envs are stepped in index order; when
    `terminated or truncated` the info gets `terminal_observation`,
    `TimeLimit.truncated = truncated and not terminated` and Monitor's
    `episode = {r, l, t}` (r = round(sum of python-float rewards, 6)), and the
    returned observation is the post-reset one.  Call sites:
    A2C_training.py:124,218; trainingCode.py:109,130,216.
    """

    def __init__(self, num_envs: int, maps: Optional[List[List[Tuple[np.ndarray, Tuple[int, int]]]]] = None,
                 curriculum: Optional[Dict[str, Any]] = None, **env_kwargs):
        self.num_envs = num_envs
        self.envs = [PlantOSOracle(**env_kwargs) for _ in range(num_envs)]
        if curriculum is not None:      # make_env_wrapper: Monitor(CurriculumWrapper(PlantOSEnv)), A2C_training.py:116-125
            self.envs = [CurriculumOracle(e, **curriculum) for e in self.envs]
        self.obs_dim = self.envs[0].obs_dim
        # maps[i] = queue of (cells, rover) consumed at each reset of env i
        self.maps = maps
        self._cursor = [0] * num_envs
        self._ep_rewards: List[List[float]] = [[] for _ in range(num_envs)]
        self._t0 = time.time()
        self.map_log: List[List[Tuple[np.ndarray, Tuple[int, int]]]] = [[] for _ in range(num_envs)]

    def _reset_one(self, i: int):
        env = self.envs[i]
        if self.maps is None:
            obs, info = env.reset()
        else:
            cells, rover = self.maps[i][self._cursor[i]]
            self._cursor[i] += 1
            obs, info = env.reset(cells, rover)
        self.map_log[i].append((env.cell_plane(), tuple(env.rover_pos)))
        self._ep_rewards[i] = []
        return obs, info

    def reset(self) -> np.ndarray:
        out = np.zeros((self.num_envs, self.obs_dim), dtype=np.float32)
        for i in range(self.num_envs):
            out[i], _ = self._reset_one(i)
        return out

    def step(self, actions: Sequence[int]):
        n = self.num_envs
        obs = np.zeros((n, self.obs_dim), dtype=np.float32)
        rewards = np.zeros(n, dtype=np.float32)
        dones = np.zeros(n, dtype=bool)
        infos: List[Dict[str, Any]] = []
        for i in range(n):
            o, r, terminated, truncated, info = self.envs[i].step(int(actions[i]))
            self._ep_rewards[i].append(float(r))
            done = terminated or truncated
            info = dict(info)
            info["TimeLimit.truncated"] = truncated and not terminated
            info["terminated"] = terminated
            info["truncated"] = truncated
            if done:
                ep = self._ep_rewards[i]
                info["episode"] = {"r": round(sum(ep), 6), "l": len(ep),
                                   "t": round(time.time() - self._t0, 6)}
                info["terminal_observation"] = o
                o, _ = self._reset_one(i)
            obs[i] = o
            rewards[i] = r
            dones[i] = done
            infos.append(info)
        return obs, rewards, dones, infos
