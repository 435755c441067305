"""Replay-harness backend that drives the CUDA simulator through PlantOSVecEnv (-> ctypes ->
the C ABI of include/plantos.h) with recorded maps injected."""
from __future__ import annotations

import numpy as np

from replay import fixture_curriculum, fixture_kwargs


class GpuBackend:
    def __init__(self, fx, kernel: str = "auto", device: str = "cuda:0", replicas: int = 1):
        """`replicas` > 1 runs that many identical copies of the fixture's envs side by side
        (env index = replica * n + i) so that a small fixture fills several 32-env pipeline
        stages of the fast kernel; every copy must behave identically and copy 0 is returned."""
        from rl_env_b200 import PlantOSVecEnv
        from rl_env_b200.vec_env import CURRICULA
        self.n = fx["actions"].shape[1]
        self.reps = replicas
        cur = fixture_curriculum(fx)
        if cur is not None:     # the reference's CurriculumWrapper variant, with the fixture's overrides
            cur = dict(CURRICULA[cur["variant"]], max_episodes_per_maze=cur["max_episodes_per_maze"])
        self.env = PlantOSVecEnv(self.n * replicas, device=device, map_source="injected", kernel=kernel,
                                 max_steps=int(fx["cfg_max_steps"]), full_infos=False, curriculum=cur,
                                 **fixture_kwargs(fx))
        self.env.push_maps(np.tile(fx["maps_cells"], (replicas, 1, 1, 1)), np.tile(fx["maps_rover"], (replicas, 1, 1)))

    def _fold(self, a):
        """[reps * n, ...] -> copy 0, after checking that all copies agree."""
        a = np.asarray(a)
        if self.reps == 1:
            return a
        r = a.reshape((self.reps, self.n) + a.shape[1:])
        assert (r == r[:1]).all(), "replicated envs diverged"
        return r[0]

    def reset(self):
        return self._fold(self.env.reset().cpu().numpy())

    def step(self, actions):
        env = self.env
        obs, rew, dones, _ = env.step(np.tile(np.asarray(actions, np.int64), self.reps))
        done = dones.cpu().numpy()
        live = {k: v.cpu().numpy() for k, v in env.scalars(False).items()}
        src = live
        ep_r = np.zeros(self.n * self.reps)
        ep_l = np.zeros(self.n * self.reps, np.int64)
        if done.any():
            term = {k: v.cpu().numpy() for k, v in env.scalars(True).items()}
            src = {k: np.where(done, term[k], live[k]) for k in live}
            ep_r = env.returns(True).cpu().numpy()
            ep_l = term["step_count"]
        state = {"x": src["x"], "y": src["y"], "step_count": src["step_count"],
                 "explored": src["explored_cells"], "total_cells": src["total_cells"],
                 "thirsty": src["thirsty_plants"], "collisions": src["total_collisions"],
                 "collided": src["collided_with_wall"]}
        f = self._fold
        return {"obs": f(obs.cpu().numpy()), "reward": f(rew.cpu().numpy()),
                "terminated": f(env.terminated.cpu().numpy()), "truncated": f(env.truncated.cpu().numpy()),
                "terminal_obs": f(env.terminal_observation.cpu().numpy()),
                "state": {k: f(v) for k, v in state.items()}, "ep_r": f(ep_r), "ep_l": f(ep_l)}

    def planes(self):
        st = self.env.get_state()
        return self._fold(st["cells"].cpu().numpy()), self._fold(st["visits"].cpu().numpy())

    def close(self):
        self.env.check()
        self.env.close()
