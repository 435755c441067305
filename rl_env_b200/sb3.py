"""Stable-Baselines3 adaptor: the batched simulator as a real `VecEnv` subclass.

The reference's trainers build `DummyVecEnv([make_env_wrapper(...)] * n)` and hand it to an SB3
algorithm (A2C_training.py:216-218,229-247; trainingCode.py:130,216).  `SB3PlantOSVecEnv` is the
drop-in for that object: SB3's numpy contract on the outside (observations float32 [N, D], rewards
float32 [N], dones bool [N], `infos` a list of dicts with `terminal_observation`,
`TimeLimit.truncated` and Monitor's `episode`), `PlantOSVecEnv` on the inside.  Actions go up and
observations / rewards / dones come down through pinned host buffers owned by the adaptor.

stable_baselines3 / gymnasium are imported lazily: `make_sb3_vecenv` raises ImportError with a clear
message when they are absent (they are not part of this repository's image); the class itself is
built by `vecenv_class(base)` so that tests can supply a stand-in base class.
"""
from __future__ import annotations

from typing import Any, List, Optional, Sequence, Type

import numpy as np


def vecenv_class(base: Type) -> Type:
    """`SB3PlantOSVecEnv` derived from `base` (stable_baselines3.common.vec_env.VecEnv)."""

    class SB3PlantOSVecEnv(base):  # type: ignore[misc, valid-type]
        """SB3 `VecEnv` over a `PlantOSVecEnv` (see module docstring)."""

        def __init__(self, env, info_mode: str = "done"):
            """`env`: a constructed PlantOSVecEnv.  `info_mode`: "done" builds info dicts only for
            envs that finished an episode (all an SB3 algorithm reads: `terminal_observation`,
            `TimeLimit.truncated`, `episode`), "full" builds the reference's 12 keys for every env
            every step (slow for large N: one device read-back per step)."""
            if info_mode not in ("done", "full"):
                raise ValueError("info_mode must be 'done' or 'full'")
            self.env = env
            self.info_mode = info_mode
            self._actions: Optional[np.ndarray] = None
            self._pinned = None
            base.__init__(self, env.num_envs, env.observation_space, env.action_space)

        # -- the three calls of the hot path
        def reset(self) -> np.ndarray:
            obs = self.env.reset()
            return obs.cpu().numpy()

        def step_async(self, actions: np.ndarray) -> None:
            self._actions = np.asarray(actions, dtype=np.int64).reshape(-1)

        def step_wait(self):
            if self._actions is None:
                raise RuntimeError("step_wait() without step_async()")
            import torch
            env = self.env
            on_gpu = getattr(env.device, "type", "cuda") == "cuda"
            if self._pinned is None:
                n, d = env.num_envs, env.obs_dim
                pin = (lambda t: t.pin_memory()) if on_gpu else (lambda t: t)
                self._pinned = {"act": pin(torch.empty(n, dtype=torch.int64)),
                                "obs": pin(torch.empty((n, d), dtype=torch.float32)),
                                "rew": pin(torch.empty(n, dtype=torch.float32)),
                                "done": pin(torch.empty(n, dtype=torch.bool))}
            pb = self._pinned
            pb["act"].numpy()[:] = self._actions
            self._actions = None
            env.step_async(pb["act"].to(env.device, non_blocking=True))
            obs_t, rew_t, done_t, lazy = env.step_wait()
            pb["obs"].copy_(obs_t, non_blocking=True)
            pb["rew"].copy_(rew_t, non_blocking=True)
            pb["done"].copy_(done_t, non_blocking=True)
            if on_gpu:
                torch.cuda.current_stream(env.device).synchronize()
            obs, rewards, dones = pb["obs"].numpy().copy(), pb["rew"].numpy().copy(), pb["done"].numpy().copy()
            if self.info_mode == "full":
                infos: List[dict] = [self._np_info(lazy[i]) for i in range(env.num_envs)]
            else:
                infos = [{} for _ in range(env.num_envs)]
                for i in np.nonzero(dones)[0]:
                    infos[int(i)] = self._np_info(lazy[int(i)])
            return obs, rewards, dones, infos

        @staticmethod
        def _np_info(info: dict) -> dict:
            tobs = info.get("terminal_observation")
            if tobs is not None and not isinstance(tobs, np.ndarray):
                info["terminal_observation"] = tobs.cpu().numpy()
            return info

        def close(self) -> None:
            self.env.close()

        # -- attribute / method plumbing SB3 expects from every VecEnv
        def get_attr(self, attr_name: str, indices=None) -> List[Any]:
            return self.env.get_attr(attr_name, indices)

        def set_attr(self, attr_name: str, value: Any, indices=None) -> None:
            self.env.set_attr(attr_name, value, indices)

        def env_method(self, method_name: str, *method_args, indices=None, **method_kwargs) -> List[Any]:
            return self.env.env_method(method_name, *method_args, indices=indices, **method_kwargs)

        def env_is_wrapped(self, wrapper_class, indices=None) -> List[bool]:
            return self.env.env_is_wrapped(wrapper_class, indices)

        def seed(self, seed: Optional[int] = None) -> Sequence[Optional[int]]:
            return self.env.seed(seed)

    return SB3PlantOSVecEnv


def make_sb3_vecenv(num_envs: int, device: Any = "cuda:0", info_mode: str = "done", **env_kwargs):
    """`DummyVecEnv([lambda: Monitor(PlantOSEnv(**kw))] * n)` replaced: builds the batched simulator and
    wraps it as an SB3 VecEnv.  `env_kwargs` are PlantOSVecEnv's (grid_size, num_plants, ..., curriculum,
    info_keywords)."""
    try:
        from stable_baselines3.common.vec_env import VecEnv  # type: ignore
    except Exception as exc:  # pragma: no cover - SB3 is not installed in the build image
        raise ImportError("stable_baselines3 is required for make_sb3_vecenv (pip install stable-baselines3); "
                          "PlantOSVecEnv itself has the same methods on torch tensors") from exc
    from .vec_env import PlantOSVecEnv
    env = PlantOSVecEnv(num_envs, device, **env_kwargs)
    return vecenv_class(VecEnv)(env, info_mode)
