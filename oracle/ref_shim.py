"""TEST INFRASTRUCTURE ONLY -- loader for the UNMODIFIED reference `plantos_env.py`.

Only usable in the build container, where the reference checkout is mounted
read-only at /root/reference.  It is used for exactly two things:

  * pinning the CPU restatement in `oracle/plantos_oracle.py` against the real
    reference (tests/test_oracle_vs_reference.py, skipped when the checkout is
    absent, e.g. on the GPU box), and
  * generating the committed golden vectors (`tests/golden/make_golden.py`).

Nothing in the product path (`rl_env_b200/`) imports this module.

The reference imports `gymnasium`, `pygame`, `plantos_3d_viewer` at module top
(plantos_env.py:1-10); none of them is installed here and none takes part in
the env arithmetic (only `math`, `random`, `numpy` do), so they are replaced
by inert stubs before the import.
"""
from __future__ import annotations

import importlib
import os
import sys
import types

import numpy as np

REFERENCE_DIR = os.environ.get("PLANTOS_REFERENCE_DIR", "/root/reference")


def reference_available() -> bool:
    return os.path.isfile(os.path.join(REFERENCE_DIR, "plantos_env.py"))


def _install_stubs() -> None:
    if "gymnasium" not in sys.modules:
        gym = types.ModuleType("gymnasium")

        class Env:  # gymnasium.Env: reset(seed=) only seeds self.np_random
            metadata: dict = {}

            def reset(self, *, seed=None, options=None):
                if seed is not None:
                    self.np_random = np.random.default_rng(seed)

            def close(self):
                pass

        class Wrapper(Env):
            def __init__(self, env):
                self.env = env

            def __getattr__(self, name):
                return getattr(self.env, name)

        class _Discrete:
            def __init__(self, n):
                self.n = int(n)
                self._rng = np.random.default_rng(0)

            def contains(self, x):
                return isinstance(x, (int, np.integer)) and 0 <= int(x) < self.n

            def sample(self):
                return int(self._rng.integers(self.n))

        class _Box:
            def __init__(self, low, high, shape, dtype):
                self.shape = tuple(shape)
                self.dtype = np.dtype(dtype)
                self.low = np.full(self.shape, low, dtype=self.dtype)
                self.high = np.full(self.shape, high, dtype=self.dtype)

        spaces = types.ModuleType("gymnasium.spaces")
        spaces.Discrete = _Discrete
        spaces.Box = _Box
        gym.Env = Env
        gym.Wrapper = Wrapper
        gym.spaces = spaces
        gym.register = lambda *a, **k: None
        sys.modules["gymnasium"] = gym
        sys.modules["gymnasium.spaces"] = spaces
    if "pygame" not in sys.modules:
        sys.modules["pygame"] = types.ModuleType("pygame")
    if "plantos_3d_viewer" not in sys.modules:
        viewer = types.ModuleType("plantos_3d_viewer")
        viewer.PlantOS3DViewer = type("PlantOS3DViewer", (), {})
        sys.modules["plantos_3d_viewer"] = viewer


def load_reference():
    """Return the reference module `plantos_env` (unmodified source)."""
    if not reference_available():
        raise RuntimeError(f"reference checkout not found at {REFERENCE_DIR}")
    _install_stubs()
    if REFERENCE_DIR not in sys.path:
        sys.path.insert(0, REFERENCE_DIR)
    return importlib.import_module("plantos_env")


class ReferenceEnv:
    """The real `PlantOSEnv` plus the written hydrated-watering policy.

    plantos_env.py:213-222: watering an already-hydrated plant falls off the
    end of `_handle_watering` (the `return self.R_MISTAKE` at :220 is dead
    code), returns None and `step` raises TypeError at :169 -- after
    `step_count` was already incremented at :162 and with no other state
    change.  The documented result (README.md:46, fixed fork
    gradio-app/plantos_env_new.py:236-245) is R_MISTAKE = -10.  This wrapper
    substitutes exactly that and then finishes `step` the way :171-183 do.
    `mistake_steps` counts how often it fired.
    """

    def __init__(self, **kwargs):
        mod = load_reference()
        self.env = mod.PlantOSEnv(**kwargs)
        self.mistake_steps = 0

    def __getattr__(self, name):
        return getattr(self.env, name)

    def reset(self, seed=None, **kwargs):
        return self.env.reset(seed=seed, **kwargs)

    # CurriculumWrapper reads and ASSIGNS env.visit_counts (A2C_training.py:85-87)
    @property
    def visit_counts(self):
        return self.env.visit_counts

    @visit_counts.setter
    def visit_counts(self, value):
        self.env.visit_counts = value

    def step(self, action):
        env = self.env
        try:
            return env.step(action)
        except TypeError:
            # only the documented failure is substituted: watering while standing on an already hydrated plant
            # (plants[pos] is False there and the reference adds `None`); any other TypeError is a real error
            if not (int(action) >= 4 and env.plants.get(env.rover_pos) is False):
                raise
            # state at this point: step_count already += 1, nothing else changed
            self.mistake_steps += 1
            reward = env.R_STEP
            reward += env.R_MISTAKE
            obs = env._get_obs()
            info = env._get_info()
            terminated = env._is_episode_done(info)
            truncated = env.step_count >= env.max_steps
            if info["exploration_percentage"] >= 100 and not env.completion_bonus_given:
                reward += env.R_COMPLETE_EXPLORATION
                env.completion_bonus_given = True
            return obs, reward, terminated, truncated, info


def load_curriculum_wrapper(variant: str = "a2c"):
    """The reference's own `CurriculumWrapper` class, executed from its source file in place
    (A2C_training.py:37-109 for "a2c", trainingCode.py:24-98 for "dqn").  The training scripts
    cannot be imported (stable_baselines3 / torch training code at module level), so only the
    class definition is compiled, against the same gymnasium stub plantos_env.py gets."""
    import ast
    if not reference_available():
        raise RuntimeError(f"reference checkout not found at {REFERENCE_DIR}")
    _install_stubs()
    path = os.path.join(REFERENCE_DIR, {"a2c": "A2C_training.py", "dqn": "trainingCode.py"}[variant])
    source = open(path).read()
    node = next(n for n in ast.parse(source).body if isinstance(n, ast.ClassDef) and n.name == "CurriculumWrapper")
    module = ast.Module(body=[node], type_ignores=[])
    ns = {"gym": sys.modules["gymnasium"], "np": np}
    exec(compile(module, path, "exec"), ns)
    return ns["CurriculumWrapper"]
