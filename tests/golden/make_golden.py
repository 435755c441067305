#!/usr/bin/env python3
"""Generate the committed golden trajectories from the UNMODIFIED reference.

Runs only in the build container (needs the read-only checkout at
/root/reference); the resulting `tests/golden/*.npz` travel with the repo.

    python tests/golden/make_golden.py            # rewrites every fixture

Each fixture records `n` reference envs stepped in `DummyVecEnv` order
(A2C_training.py:216-218 -- index order, auto-reset on done) with the global
`random` module seeded once (maps come from it, plantos_env.py:344-372) and
actions from `numpy.random.default_rng(seed)`.  Hydrated-plant watering uses
the documented -10 substitution (see oracle/ref_shim.py).

Arrays (T steps, n envs, D obs dim, G grid, E = max episodes seen per env):
  cfg_*            ctor kwargs, max_steps, seed
  maps_cells       u8  [n,E,G,G]  cell codes 0 empty 1 obstacle 2 hydrated 3 thirsty
  maps_rover       i16 [n,E,2]    rover start (x,y); n_maps i32 [n] valid count
  reset_obs        f32 [n,D]      observations returned by the initial reset
  actions          i64 [T,n]
  obs              f32 [T,n,D]    VecEnv observation (post-reset where done)
  rewards          f64 [T,n]      python-float rewards, exact
  terminated, truncated  bool [T,n]
  term_t, term_i   i32 [K]; term_obs f32 [K,D]  terminal observations; ep_r f64 [K], ep_l i32 [K]
  x,y,step_count,explored,total_cells,thirsty,collisions,collided  i32 [T,n]
                   info of the step (pre-reset), plantos_env.py:317-336
  lidar_dist, lidar_kind  u8 [T,n,C]  integer LIDAR hit (pre-reset)
  visit_hash       u64 [T,n]      hash of visit_counts pre-reset (see visit_hash())
  snap_t i32 [S]; snap_visits i32 [S,n,G,G]; snap_cells u8 [S,n,G,G]
                   full planes AFTER step snap_t (post-reset where done)
"""
from __future__ import annotations

import os
import random
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)

from oracle.ref_shim import ReferenceEnv, load_curriculum_wrapper  # noqa: E402

OUT = os.path.dirname(os.path.abspath(__file__))


def visit_hash(v: np.ndarray) -> np.uint64:
    """Order-sensitive 64-bit hash of an int visit plane [G,G] (wraps mod 2^64)."""
    flat = np.asarray(v).reshape(-1).astype(np.uint64)
    w = (np.arange(flat.size, dtype=np.uint64) * np.uint64(2654435761) + np.uint64(1))
    with np.errstate(over="ignore"):
        return np.uint64((flat * w).sum(dtype=np.uint64))


def cell_plane(env) -> np.ndarray:
    g = env.grid_size
    p = np.zeros((g, g), dtype=np.uint8)
    for (x, y) in env.obstacles:
        p[x, y] = 1
    for (x, y), thirsty in env.plants.items():
        p[x, y] = 3 if thirsty else 2
    return p


def lidar_ints(env, obs: np.ndarray):
    c, r = env.lidar_channels, env.lidar_range
    lid = obs[: 5 * c].reshape(c, 5)
    dist = np.rint(lid[:, 0] * r).astype(np.uint8)
    kind = np.argmax(lid[:, 1:], axis=1).astype(np.uint8)
    return dist, kind


class _Raw:
    """Uniform access for `record`: `.env` is the raw PlantOSEnv whether or not the reference's
    CurriculumWrapper sits in between; reset/step go through the outermost layer."""

    def __init__(self, outer, raw_holder):
        self.outer, self.holder = outer, raw_holder

    @property
    def env(self):
        return self.holder.env

    @property
    def mistake_steps(self):
        return self.holder.mistake_steps

    def reset(self):
        return self.outer.reset()

    def step(self, a):
        return self.outer.step(a)


def record(name: str, n: int, steps: int, seed: int, kwargs: dict, max_steps: int = 1000,
           snap_every: int = 250, curriculum: str = "", cur_max_eps: int = 0) -> None:
    """`curriculum` = "a2c" / "dqn": every env is wrapped in the reference's OWN CurriculumWrapper
    (class compiled from A2C_training.py / trainingCode.py, constructed as at their call sites
    :121 / :107) -- Monitor(CurriculumWrapper(PlantOSEnv)) as in make_env_wrapper.
    `cur_max_eps` overrides max_episodes_per_maze (a plain attribute) so that short fixtures see
    the time-out branch of the 50-episode variant too."""
    random.seed(seed)
    np.random.seed(seed)   # CurriculumWrapper draws (unused) maze seeds from numpy's global RNG
    rng = np.random.default_rng(seed)
    holders = [ReferenceEnv(**kwargs) for _ in range(n)]
    if curriculum:
        wrapper = load_curriculum_wrapper(curriculum)
        init = {"a2c": 40.0, "dqn": 30.0}[curriculum]
        outers = [wrapper(h, initial_threshold=init, max_threshold=100.0) for h in holders]
        for o in outers:
            if cur_max_eps:
                o.max_episodes_per_maze = cur_max_eps
    else:
        outers = holders
    envs = [_Raw(o, h) for o, h in zip(outers, holders)]
    for e in envs:
        e.env.max_steps = max_steps  # plain attribute, plantos_env.py:120
    g = kwargs["grid_size"]
    c = kwargs["lidar_channels"]
    d = 5 * c + 27
    maps = [[] for _ in range(n)]
    reset_obs = np.zeros((n, d), np.float32)
    for i, e in enumerate(envs):
        o, _ = e.reset()
        maps[i].append((cell_plane(e.env), e.env.rover_pos))
        reset_obs[i] = o
    actions = rng.integers(0, 5, size=(steps, n)).astype(np.int64)
    obs = np.zeros((steps, n, d), np.float32)
    rewards = np.zeros((steps, n), np.float64)
    terminated = np.zeros((steps, n), bool)
    truncated = np.zeros((steps, n), bool)
    ints = {k: np.zeros((steps, n), np.int32) for k in
            ("x", "y", "step_count", "explored", "total_cells", "thirsty", "collisions", "collided")}
    lidar_dist = np.zeros((steps, n, c), np.uint8)
    lidar_kind = np.zeros((steps, n, c), np.uint8)
    vhash = np.zeros((steps, n), np.uint64)
    cur_thr = np.zeros((steps, n), np.float64)
    term_t, term_i, term_obs, ep_r, ep_l = [], [], [], [], []
    ep_rewards = [[] for _ in range(n)]
    snap_t, snap_visits, snap_cells = [], [], []
    mistakes = 0
    for t in range(steps):
        for i, e in enumerate(envs):
            o, r, te, tr, info = e.step(int(actions[t, i]))
            ep_rewards[i].append(float(r))
            rewards[t, i] = r
            terminated[t, i] = te
            truncated[t, i] = tr
            ints["x"][t, i], ints["y"][t, i] = info["rover_position"]
            ints["step_count"][t, i] = info["step_count"]
            ints["explored"][t, i] = info["explored_cells"]
            ints["total_cells"][t, i] = info["total_cells"]
            ints["thirsty"][t, i] = info["thirsty_plants"]
            ints["collisions"][t, i] = info["total_collisions"]
            ints["collided"][t, i] = info["collided_with_wall"]
            lidar_dist[t, i], lidar_kind[t, i] = lidar_ints(e.env, o)
            vhash[t, i] = visit_hash(e.env.visit_counts)
            if curriculum:
                cur_thr[t, i] = e.outer.exploration_threshold
            if te or tr:
                term_t.append(t)
                term_i.append(i)
                term_obs.append(o)
                ep_r.append(round(sum(ep_rewards[i]), 6))  # SB3 Monitor's "r"
                ep_l.append(len(ep_rewards[i]))
                ep_rewards[i] = []
                o, _ = e.reset()
                maps[i].append((cell_plane(e.env), e.env.rover_pos))
            obs[t, i] = o
        if (t + 1) % snap_every == 0 or t == steps - 1:
            snap_t.append(t)
            snap_visits.append(np.stack([e.env.visit_counts.astype(np.int32) for e in envs]))
            snap_cells.append(np.stack([cell_plane(e.env) for e in envs]))
    mistakes = sum(e.mistake_steps for e in envs)
    emax = max(len(m) for m in maps)
    maps_cells = np.zeros((n, emax, g, g), np.uint8)
    maps_rover = np.zeros((n, emax, 2), np.int16)
    n_maps = np.zeros(n, np.int32)
    for i, m in enumerate(maps):
        n_maps[i] = len(m)
        for k, (cells, rover) in enumerate(m):
            maps_cells[i, k] = cells
            maps_rover[i, k] = rover
    path = os.path.join(OUT, name + ".npz")
    np.savez_compressed(
        path,
        cfg_grid_size=g, cfg_num_plants=kwargs["num_plants"], cfg_num_obstacles=kwargs["num_obstacles"],
        cfg_lidar_range=kwargs["lidar_range"], cfg_lidar_channels=c, cfg_max_steps=max_steps,
        cfg_seed=seed, cfg_mistake_steps=mistakes, cfg_curriculum=curriculum,
        cfg_cur_max_eps=(cur_max_eps or {"": 0, "a2c": 3, "dqn": 50}[curriculum]), cur_threshold=cur_thr,
        maps_cells=maps_cells, maps_rover=maps_rover, n_maps=n_maps, reset_obs=reset_obs,
        actions=actions, obs=obs, rewards=rewards, terminated=terminated, truncated=truncated,
        term_t=np.array(term_t, np.int32), term_i=np.array(term_i, np.int32),
        term_obs=np.array(term_obs, np.float32).reshape(len(term_t), d),
        ep_r=np.array(ep_r, np.float64), ep_l=np.array(ep_l, np.int32),
        lidar_dist=lidar_dist, lidar_kind=lidar_kind, visit_hash=vhash,
        snap_t=np.array(snap_t, np.int32), snap_visits=np.array(snap_visits, np.int32),
        snap_cells=np.array(snap_cells, np.uint8), **ints)
    print(f"{name}: n={n} T={steps} episodes={len(term_t)} terminated={int(terminated.sum())} "
          f"mistake_steps={mistakes} -> {os.path.getsize(path) / 1024:.0f} KiB")


FIXTURES = [
    # BASELINE.json configs[1]: 8-env replay, training preset (A2C_training.py:206-212)
    ("replay_T_8env", 8, 3000, 1234,
     dict(grid_size=25, num_plants=10, num_obstacles=12, lidar_range=6, lidar_channels=16), 1000),
    # BASELINE.json configs[0]: ctor-default single env (README.md:99-122)
    ("replay_DFLT_1env", 1, 2000, 0,
     dict(grid_size=21, num_plants=8, num_obstacles=50, lidar_range=2, lidar_channels=10), 1000),
    # tiny grids: rays longer than the grid, full exploration -> terminated + completion bonus
    ("replay_tiny_4env", 4, 1500, 7,
     dict(grid_size=7, num_plants=3, num_obstacles=3, lidar_range=8, lidar_channels=8), 1000),
    # odd ray count, truncation-heavy (max_steps override)
    ("replay_odd_4env", 4, 600, 11,
     dict(grid_size=12, num_plants=5, num_obstacles=9, lidar_range=4, lidar_channels=5), 50),
    # XL stress preset, reset-heavy
    ("replay_XL_2env", 2, 250, 5,
     dict(grid_size=64, num_plants=64, num_obstacles=600, lidar_range=32, lidar_channels=16), 100),
]

# CurriculumWrapper fixtures: small grids so that thresholds are reached and raised
CURRICULUM_FIXTURES = [
    ("replay_curr_a2c_4env", 4, 3000, 21,
     dict(grid_size=6, num_plants=2, num_obstacles=3, lidar_range=4, lidar_channels=8), 150, "a2c", 0),   # (R, C) with a fast-kernel instantiation
    ("replay_curr_dqn_3env", 3, 3000, 22,
     dict(grid_size=6, num_plants=2, num_obstacles=3, lidar_range=4, lidar_channels=6), 200, "dqn", 4),
]

if __name__ == "__main__":
    only = set(sys.argv[1:])
    for name, n, steps, seed, kw, ms in FIXTURES:
        if only and name not in only:
            continue
        record(name, n, steps, seed, kw, max_steps=ms)
    for name, n, steps, seed, kw, ms, cur, cme in CURRICULUM_FIXTURES:
        if only and name not in only:
            continue
        record(name, n, steps, seed, kw, max_steps=ms, curriculum=cur, cur_max_eps=cme)
