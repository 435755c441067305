"""Replay-harness backend that drives the CUDA simulator through PlantOSVecEnv (-> ctypes ->
the C ABI of include/plantos.h) with recorded maps injected."""
from __future__ import annotations

import numpy as np

from replay import fixture_kwargs


class GpuBackend:
    def __init__(self, fx, kernel: str = "auto", device: str = "cuda:0"):
        from rl_env_b200 import PlantOSVecEnv
        self.n = fx["actions"].shape[1]
        self.env = PlantOSVecEnv(self.n, device=device, map_source="injected", kernel=kernel,
                                 max_steps=int(fx["cfg_max_steps"]), full_infos=False, **fixture_kwargs(fx))
        self.env.push_maps(fx["maps_cells"], fx["maps_rover"])

    def reset(self):
        return self.env.reset().cpu().numpy()

    def step(self, actions):
        env = self.env
        obs, rew, dones, _ = env.step(np.asarray(actions, np.int64))
        done = dones.cpu().numpy()
        live = {k: v.cpu().numpy() for k, v in env.scalars(False).items()}
        src = live
        ep_r = np.zeros(self.n)
        ep_l = np.zeros(self.n, np.int64)
        if done.any():
            term = {k: v.cpu().numpy() for k, v in env.scalars(True).items()}
            src = {k: np.where(done, term[k], live[k]) for k in live}
            ep_r = env.returns(True).cpu().numpy()
            ep_l = term["step_count"]
        state = {"x": src["x"], "y": src["y"], "step_count": src["step_count"],
                 "explored": src["explored_cells"], "total_cells": src["total_cells"],
                 "thirsty": src["thirsty_plants"], "collisions": src["total_collisions"],
                 "collided": src["collided_with_wall"]}
        return {"obs": obs.cpu().numpy(), "reward": rew.cpu().numpy(),
                "terminated": env.terminated.cpu().numpy(), "truncated": env.truncated.cpu().numpy(),
                "terminal_obs": env.terminal_observation.cpu().numpy(), "state": state,
                "ep_r": ep_r, "ep_l": ep_l}

    def planes(self):
        st = self.env.get_state()
        return st["cells"].cpu().numpy(), st["visits"].cpu().numpy()

    def close(self):
        self.env.check()
        self.env.close()
