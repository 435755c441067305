"""Shared replay harness: run a recorded reference trajectory (tests/golden/*.npz) through a
backend and compare every step.  Backends: the Python oracle port, the C oracle, and the
CUDA simulator (through PlantOSVecEnv -> C ABI)."""
from __future__ import annotations

import os
import sys
from typing import Dict, List

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

GOLDEN_DIR = os.path.join(ROOT, "tests", "golden")
FIXTURES = ["replay_T_8env", "replay_DFLT_1env", "replay_tiny_4env", "replay_odd_4env", "replay_XL_2env"]
CURRICULUM_FIXTURES = ["replay_curr_a2c_4env", "replay_curr_dqn_3env"]   # recorded through the reference's CurriculumWrapper
STATE_KEYS = ("x", "y", "step_count", "explored", "total_cells", "thirsty", "collisions", "collided")
FLOAT_TOL = 1e-5  # BASELINE.json north_star: rewards / float observations within 1e-5 absolute


def load_fixture(name: str) -> Dict[str, np.ndarray]:
    with np.load(os.path.join(GOLDEN_DIR, name + ".npz")) as z:
        return {k: z[k] for k in z.files}


def fixture_kwargs(fx) -> Dict[str, int]:
    return dict(grid_size=int(fx["cfg_grid_size"]), num_plants=int(fx["cfg_num_plants"]),
                num_obstacles=int(fx["cfg_num_obstacles"]), lidar_range=int(fx["cfg_lidar_range"]),
                lidar_channels=int(fx["cfg_lidar_channels"]))


def fixture_curriculum(fx):
    """None, or the CurriculumWrapper variant + overrides the fixture was recorded with."""
    variant = str(fx["cfg_curriculum"]) if "cfg_curriculum" in fx else ""
    if not variant:
        return None
    return dict(variant=variant, max_episodes_per_maze=int(fx["cfg_cur_max_eps"]))


def visit_hash(v: np.ndarray) -> np.uint64:
    """Same hash as tests/golden/make_golden.py."""
    flat = np.asarray(v).reshape(-1).astype(np.uint64)
    w = (np.arange(flat.size, dtype=np.uint64) * np.uint64(2654435761) + np.uint64(1))
    with np.errstate(over="ignore"):
        return np.uint64((flat * w).sum(dtype=np.uint64))


def lidar_ints(obs: np.ndarray, channels: int, rng: int):
    lid = obs[..., : 5 * channels].reshape(obs.shape[:-1] + (channels, 5))
    dist = np.rint(lid[..., 0] * rng).astype(np.uint8)
    kind = np.argmax(lid[..., 1:], axis=-1).astype(np.uint8)
    return dist, kind


class Mismatch(AssertionError):
    pass


def _eq(name, t, got, want, exact=True):
    got, want = np.asarray(got), np.asarray(want)
    if exact:
        ok = np.array_equal(got, want)
    else:
        ok = got.shape == want.shape and np.all(np.abs(got.astype(np.float64) - want.astype(np.float64)) <= FLOAT_TOL)
    if not ok:
        bad = np.argwhere(np.asarray(got != want)) if got.shape == want.shape else None
        first = tuple(bad[0]) if bad is not None and len(bad) else None
        detail = ""
        if first is not None:
            detail = f" first at {first}: got {got[first]!r} want {want[first]!r}; {len(bad)} entries differ"
        raise Mismatch(f"step {t}: {name} differs.{detail}")


def check_replay(fx, backend, steps: int = None, check_state_every: int = 1) -> Dict[str, int]:
    """Drive `backend` with the fixture's maps and actions; raise Mismatch on any difference.

    backend API:
      reset() -> obs [n,D] f32
      step(actions i64 [n]) -> dict(obs, reward, terminated, truncated, terminal_obs [n,D],
                                    state {STATE_KEYS -> int [n]} as of the step, pre-reset,
                                    ep_r [n] f64, ep_l [n] int  (valid where done))
      planes() -> (cells u8 [n,G,G], visits i32 [n,G,G])  current (post-reset) planes
    Returns counters (steps, episodes, bit-exact float stats).
    """
    c, r = int(fx["cfg_lidar_channels"]), int(fx["cfg_lidar_range"])
    T = fx["actions"].shape[0] if steps is None else min(steps, fx["actions"].shape[0])
    obs0 = backend.reset()
    _eq("reset obs", -1, obs0, fx["reset_obs"], exact=False)
    bitexact_obs = int(np.array_equal(obs0.view(np.uint32), fx["reset_obs"].view(np.uint32)))
    bitexact_rew = 1
    term_lookup = {(int(t), int(i)): k for k, (t, i) in enumerate(zip(fx["term_t"], fx["term_i"]))}
    snap_lookup = {int(t): k for k, t in enumerate(fx["snap_t"])}
    episodes = 0
    for t in range(T):
        out = backend.step(fx["actions"][t])
        done = fx["terminated"][t] | fx["truncated"][t]
        _eq("terminated", t, np.asarray(out["terminated"], bool), fx["terminated"][t])
        _eq("truncated", t, np.asarray(out["truncated"], bool), fx["truncated"][t])
        _eq("obs", t, out["obs"], fx["obs"][t], exact=False)
        bitexact_obs &= int(np.array_equal(np.asarray(out["obs"]).view(np.uint32), fx["obs"][t].view(np.uint32)))
        want_r32 = fx["rewards"][t].astype(np.float32)
        _eq("reward", t, np.asarray(out["reward"], np.float32), want_r32, exact=False)
        bitexact_rew &= int(np.array_equal(np.asarray(out["reward"], np.float32).view(np.uint32), want_r32.view(np.uint32)))
        for key in STATE_KEYS:
            _eq("state." + key, t, out["state"][key], fx[key][t])
        # integer LIDAR hits of the pre-reset observation
        pre = np.array(out["obs"], copy=True)
        for i in np.nonzero(done)[0]:
            k = term_lookup[(t, int(i))]
            _eq(f"terminal_obs[env {i}]", t, out["terminal_obs"][i], fx["term_obs"][k], exact=False)
            pre[i] = out["terminal_obs"][i]
            if "ep_r" in out:
                if round(float(out["ep_r"][i]), 6) != float(fx["ep_r"][k]):
                    raise Mismatch(f"step {t}: episode return env {i}: {out['ep_r'][i]!r} vs {fx['ep_r'][k]!r}")
                _eq(f"episode length[env {i}]", t, int(out["ep_l"][i]), int(fx["ep_l"][k]))
            episodes += 1
        dist, kind = lidar_ints(pre, c, r)
        _eq("lidar distance", t, dist, fx["lidar_dist"][t])
        _eq("lidar kind", t, kind, fx["lidar_kind"][t])
        if check_state_every and (t % check_state_every == 0 or t in snap_lookup):
            cells, visits = backend.planes()
            for i in np.nonzero(~done)[0]:
                if visit_hash(visits[i]) != fx["visit_hash"][t, i]:
                    raise Mismatch(f"step {t}: visit_counts of env {i} differ from the reference")
            if t in snap_lookup:
                k = snap_lookup[t]
                _eq("visit_counts snapshot", t, visits, fx["snap_visits"][k])
                _eq("cell plane snapshot", t, cells, fx["snap_cells"][k])
    return {"steps": T, "episodes": episodes, "bitexact_obs": bitexact_obs, "bitexact_reward": bitexact_rew}


# ------------------------------------------------------------------ CPU backends
class PyOracleBackend:
    def __init__(self, fx):
        from oracle.plantos_oracle import OracleVecEnv
        n = fx["actions"].shape[1]
        maps = [[(fx["maps_cells"][i, k], tuple(fx["maps_rover"][i, k])) for k in range(int(fx["n_maps"][i]))]
                for i in range(n)]
        self.env = OracleVecEnv(n, maps=maps, curriculum=fixture_curriculum(fx), max_steps=int(fx["cfg_max_steps"]),
                                **fixture_kwargs(fx))
        self.n = n

    def reset(self):
        return self.env.reset()

    def step(self, actions):
        obs, rew, dones, infos = self.env.step(actions)
        n = self.n
        term_obs = np.zeros_like(obs)
        ep_r, ep_l = np.zeros(n), np.zeros(n, np.int64)
        for i, info in enumerate(infos):
            if dones[i]:
                term_obs[i] = info["terminal_observation"]
                ep_r[i], ep_l[i] = info["episode"]["r"], info["episode"]["l"]
        state = {
            "x": [inf["rover_position"][0] for inf in infos], "y": [inf["rover_position"][1] for inf in infos],
            "step_count": [inf["step_count"] for inf in infos], "explored": [inf["explored_cells"] for inf in infos],
            "total_cells": [inf["total_cells"] for inf in infos], "thirsty": [inf["thirsty_plants"] for inf in infos],
            "collisions": [inf["total_collisions"] for inf in infos],
            "collided": [int(inf["collided_with_wall"]) for inf in infos],
        }
        return {"obs": obs, "reward": rew, "terminated": [inf["terminated"] for inf in infos],
                "truncated": [inf["truncated"] for inf in infos], "terminal_obs": term_obs,
                "state": state, "ep_r": ep_r, "ep_l": ep_l}

    def planes(self):
        cells = np.stack([e.cell_plane() for e in self.env.envs])
        visits = np.stack([e.visit_counts for e in self.env.envs]).astype(np.int32)
        return cells, visits


class COracleBackend:
    def __init__(self, fx):
        from oracle.c_oracle import COracle
        kw = fixture_kwargs(fx)
        n = fx["actions"].shape[1]
        self.o = COracle(n, kw["grid_size"], kw["num_plants"], kw["num_obstacles"], kw["lidar_range"],
                         kw["lidar_channels"], int(fx["cfg_max_steps"]))
        self.o.set_maps(fx["maps_cells"], fx["maps_rover"])
        self.n = n
        self._sc = None

    def reset(self):
        return self.o.reset().copy()

    def step(self, actions):
        obs, rew, term, trunc = self.o.step(actions)
        _, _, sc = self.o.get_state()
        done = term | trunc
        names = {"x": 0, "y": 1, "step_count": 2, "explored": 3, "total_cells": 4, "thirsty": 5,
                 "collisions": 6, "collided": 7}
        # live scalars, except where the env just finished: there the pre-reset snapshot
        state = {key: np.where(done, self.o.term_sc[k], sc[k]) for key, k in names.items()}
        out = {"obs": obs.copy(), "reward": rew.copy(), "terminated": term, "truncated": trunc,
               "terminal_obs": self.o.terminal_obs.copy(), "state": state,
               "ep_r": self.o.ep_return.copy(), "ep_l": self.o.ep_len.copy()}
        return out

    def planes(self):
        cells, visits, _ = self.o.get_state()
        return cells, visits
