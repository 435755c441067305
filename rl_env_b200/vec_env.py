"""PlantOSVecEnv -- the SB3-style `VecEnv` face of the B200 simulator.

Drop-in for `DummyVecEnv([lambda: Monitor(PlantOSEnv(**kw))] * n)` as the reference's
trainers build it (A2C_training.py:116-125,216-218; trainingCode.py:109,130,216):
`reset()`, `step_async(actions)`, `step_wait()`, `step()`, `close()`, `num_envs`,
`observation_space`, `action_space`, `get_attr`/`set_attr`/`env_method`/`env_is_wrapped`/
`seed`.  Differences a caller sees: arrays are CUDA-resident torch tensors that the next
step overwrites, and `infos` is a lazy sequence (dicts are built on access) unless
`full_infos=True`.

Everything is computed by libplantos_b200.so (include/plantos.h) -- there is no CPU path.
"""
from __future__ import annotations

import ctypes as C
import os
import time
from typing import Any, Dict, List, Optional, Sequence

import numpy as np
import torch

from . import _native as nat
from . import tables

try:  # real spaces when gymnasium is installed (it is not in the build image)
    from gymnasium import spaces as _spaces  # type: ignore
except Exception:  # pragma: no cover - exercised in the build image
    _spaces = None


class _Discrete:
    def __init__(self, n: int):
        self.n = int(n)
        self.shape = ()
        self.dtype = np.dtype(np.int64)

    def contains(self, x) -> bool:
        return isinstance(x, (int, np.integer)) and 0 <= int(x) < self.n

    def sample(self) -> int:
        return int(np.random.randint(self.n))


class _Box:
    def __init__(self, low, high, shape, dtype):
        self.shape = tuple(shape)
        self.dtype = np.dtype(dtype)
        self.low = np.full(self.shape, low, dtype=self.dtype)
        self.high = np.full(self.shape, high, dtype=self.dtype)


# the two CurriculumWrapper variants the reference ships, with the arguments of their call sites
CURRICULA: Dict[str, Dict[str, Any]] = {
    "a2c": dict(mode="terminate", initial_threshold=40.0, max_threshold=100.0, threshold_increment=10.0,
                max_episodes_per_maze=3),      # A2C_training.py:40,55,121
    "dqn": dict(mode="mark", initial_threshold=30.0, max_threshold=100.0, threshold_increment=5.0,
                max_episodes_per_maze=50),     # trainingCode.py:29,44,107
}

PRESETS: Dict[str, Dict[str, int]] = {
    # ctor defaults, plantos_env.py:25-26 (D = 77)
    "default": dict(grid_size=21, num_plants=8, num_obstacles=50, lidar_range=2, lidar_channels=10),
    # A2C_training.py:206-212 / trainingCode.py:120-126 (D = 107)
    "training": dict(grid_size=25, num_plants=10, num_obstacles=12, lidar_range=6, lidar_channels=16),
    # large-map stress configuration of BASELINE.json configs[4]
    "xl": dict(grid_size=64, num_plants=64, num_obstacles=600, lidar_range=32, lidar_channels=16),
}

_KERNELS = {"auto": nat.KERNEL_AUTO, "generic": nat.KERNEL_GENERIC, "fast": nat.KERNEL_FAST}


class LazyInfos(Sequence):
    """`infos` of one step: behaves like SB3's list of dicts, built on access.

    Keys per env: the 12 of PlantOSEnv._get_info (plantos_env.py:323-336); for finished
    envs additionally `terminal_observation`, `TimeLimit.truncated` and Monitor's
    `episode = {r, l, t}`.  The scalar snapshot is pulled from the device on first access,
    so read it before the next `step_async`.
    """

    def __init__(self, env: "PlantOSVecEnv"):
        self._env = env
        self._snap: Optional[Dict[str, np.ndarray]] = None

    def __len__(self) -> int:
        return self._env.num_envs

    def snapshot(self) -> Dict[str, np.ndarray]:
        if self._snap is None:
            self._snap = self._env._info_snapshot()
        return self._snap

    def __getitem__(self, i):
        if isinstance(i, slice):
            return [self[k] for k in range(*i.indices(len(self)))]
        n = len(self)
        if i < 0:
            i += n
        if not 0 <= i < n:
            raise IndexError(i)
        s = self.snapshot()
        env = self._env
        done = bool(s["done"][i])
        src = s["term"] if done else s["live"]
        explored, total = int(src["explored_cells"][i]), int(src["total_cells"][i])
        thirsty = int(src["thirsty_plants"][i])
        info: Dict[str, Any] = {
            "rover_position": (int(src["x"][i]), int(src["y"][i])),
            "thirsty_plants": thirsty,
            "hydrated_plants": env.num_plants - thirsty,
            "total_plants": env.num_plants,
            "step_count": int(src["step_count"][i]),
            "explored_cells": explored,
            "total_cells": total,
            "exploration_percentage": (explored / total) * 100,
            "lidar_range": env.lidar_range,
            "lidar_channels": env.lidar_channels,
            "collided_with_wall": bool(src["collided_with_wall"][i]),
            "total_collisions": int(src["total_collisions"][i]),
            "TimeLimit.truncated": bool(s["truncated"][i]) and not bool(s["terminated"][i]),
        }
        if s.get("actions") is not None:                   # the Gradio fork's flag (gradio-app/plantos_env_new.py:184)
            info["is_watering"] = bool(s["actions"][i] >= 4)
        if done:
            info["episode"] = {"r": round(float(s["term_return"][i]), 6),
                               "l": int(src["step_count"][i]),
                               "t": round(time.time() - env._t_start, 6)}
            for key in env.info_keywords:                  # SB3 Monitor: ep_info[key] = info[key]
                info["episode"][key] = info[key]
            if env._terminal_obs is not None:
                info["terminal_observation"] = env._terminal_obs[i]
        return info


class PlantOSVecEnv:
    """N PlantOS envs on one B200, one kernel launch per `step`."""

    def __init__(self, num_envs: int, device: Any = "cuda:0", *,
                 grid_size: int = 21, num_plants: int = 8, num_obstacles: int = 50,
                 lidar_range: int = 2, lidar_channels: int = 10, thirsty_plant_prob: float = 0.7,
                 max_steps: int = 1000, seed: int = 0, map_source: str = "philox",
                 env_id_base: int = 0, kernel: str = "auto",
                 rewards: Optional[Dict[str, float]] = None,
                 track_terminal_obs: bool = True, full_infos: Optional[bool] = None,
                 obs_ring: int = 1, curriculum: Any = None, info_keywords: Sequence[str] = (),
                 tuning: Optional[Dict[str, int]] = None):
        self.device = torch.device(device)
        if self.device.type != "cuda":
            raise ValueError("PlantOSVecEnv runs on a CUDA device only (no CPU path)")
        if not torch.cuda.is_available():
            raise RuntimeError("PlantOSVecEnv needs a CUDA device; torch.cuda.is_available() is False")
        if map_source not in ("philox", "injected", "maze"):
            raise ValueError("map_source must be 'philox', 'injected' or 'maze'")
        self._lib = nat.load()
        self.num_envs = int(num_envs)
        self.grid_size, self.num_plants, self.num_obstacles = int(grid_size), int(num_plants), int(num_obstacles)
        self.lidar_range, self.lidar_channels = int(lidar_range), int(lidar_channels)
        self.thirsty_plant_prob = float(thirsty_plant_prob)
        self.max_steps = int(max_steps)
        self.rewards = dict(tables.DEFAULT_REWARDS)
        if rewards:
            self.rewards.update(rewards)
        self.obs_dim = 5 * self.lidar_channels + 2 + 25
        self.map_source = map_source
        self.full_infos = (self.num_envs <= 64) if full_infos is None else bool(full_infos)

        cfg = nat.Config()
        nat.check(self._lib.plantos_default_config(C.byref(cfg)))
        cfg.num_envs = self.num_envs
        cfg.env_id_base = int(env_id_base)
        self.env_id_base = int(env_id_base)
        cfg.grid_size, cfg.num_plants, cfg.num_obstacles = self.grid_size, self.num_plants, self.num_obstacles
        cfg.lidar_range, cfg.lidar_channels = self.lidar_range, self.lidar_channels
        cfg.max_steps = self.max_steps
        cfg.thirsty_plant_prob = self.thirsty_plant_prob
        # "maze": the Gradio fork's maze generator (gradio-app/plantos_env_new.py:408-604) on the device
        cfg.map_source = {"philox": nat.MAPS_PHILOX, "injected": nat.MAPS_INJECTED, "maze": nat.MAPS_MAZE}[map_source]
        cfg.seed = int(seed) & 0xFFFFFFFFFFFFFFFF
        for key, val in self.rewards.items():
            setattr(cfg, key, float(val))
        cfg.kernel = _KERNELS[kernel]
        # tuning / test knobs of plantos_config_t (never change results): fast_grid, fast_impl ("tile" | "trip" |
        # "lane"), no_pdl, l2_keep_mb.  The PLANTOS_FAST_GRID / PLANTOS_FAST_IMPL / PLANTOS_PDL / PLANTOS_L2_PERSIST_MB
        # environment variables of round 1's experiments are honoured HERE (the C library reads no environment).
        tune = {"fast_grid": int(os.environ.get("PLANTOS_FAST_GRID", "0") or 0),
                "fast_impl": os.environ.get("PLANTOS_FAST_IMPL", "tile") or "tile",
                "no_pdl": 1 if os.environ.get("PLANTOS_PDL", "1") == "0" else 0,
                "l2_keep_mb": int(os.environ.get("PLANTOS_L2_PERSIST_MB", "0") or 0) if os.environ.get("PLANTOS_L2_KEEP", "0") == "1" else 0}
        tune.update(tuning or {})
        cfg.tune_fast_grid = int(tune["fast_grid"])
        cfg.tune_fast_impl = {"tile": 0, "trip": 1, "lane": 2}.get(tune["fast_impl"], tune["fast_impl"]) if isinstance(tune["fast_impl"], str) else int(tune["fast_impl"])
        cfg.tune_no_pdl = int(tune["no_pdl"])
        cfg.tune_l2_keep_mb = int(tune["l2_keep_mb"])
        self._cfg = cfg
        dev_index = self.device.index if self.device.index is not None else torch.cuda.current_device()
        self.device = torch.device("cuda", dev_index)
        handle = C.c_void_p()
        nat.check(self._lib.plantos_create(C.byref(cfg), dev_index, C.byref(handle)))
        self._h = handle

        # CurriculumWrapper on the device: "a2c" (A2C_training.py:37-109, :121), "dqn"
        # (trainingCode.py:24-98, :107) or a dict(mode="terminate"|"mark", initial_threshold,
        # max_threshold, threshold_increment, max_episodes_per_maze)
        self.curriculum = None
        if curriculum is not None:
            cur = dict(CURRICULA[curriculum]) if isinstance(curriculum, str) else dict(curriculum)
            base = dict(CURRICULA["a2c" if cur.get("mode", "terminate") == "terminate" else "dqn"])
            base.update(cur)
            mode = {"terminate": nat.CURRICULUM_TERMINATE, "mark": nat.CURRICULUM_MARK}[base["mode"]]
            nat.check(self._lib.plantos_set_curriculum(
                self._h, mode, float(base["initial_threshold"]), float(base["max_threshold"]),
                float(base["threshold_increment"]), int(base["max_episodes_per_maze"])))
            # reuse_map=True: a kept maze really is regenerated identically, as the wrapper's
            # `reset(seed=self.current_maze_seed)` means it to be (A2C_training.py:75-86); the
            # reference itself draws a new map every time (default)
            if base.get("reuse_map"):
                nat.check(self._lib.plantos_set_curriculum_reuse_map(self._h, 1))
            self.curriculum = base

        # tables evaluated by the Python interpreter, exactly as the reference evaluates them
        off = np.ascontiguousarray(tables.lidar_offsets(self.lidar_channels, self.lidar_range))
        dist = tables.distance_table(self.lidar_range)
        pos = tables.position_table(self.grid_size)
        vis = tables.visit_table()
        rw = tables.reward_table(self.rewards)
        nat.check(self._lib.plantos_upload_tables(
            self._h, off.ctypes.data, dist.ctypes.data, pos.ctypes.data, vis.ctypes.data, rw.ctypes.data))

        n, d, dev = self.num_envs, self.obs_dim, self.device
        # obs_ring > 1 rotates the observation output through that many buffers (one per step),
        # the way a rollout buffer [n_steps, N, D] is filled: step t's tensor stays valid for
        # obs_ring - 1 further steps instead of being overwritten by the next one.
        self._obs_ring = [torch.empty((n, d), dtype=torch.float32, device=dev) for _ in range(max(1, int(obs_ring)))]
        self._ring_pos = 0
        self._obs = self._obs_ring[0]
        self._rewards = torch.empty(n, dtype=torch.float32, device=dev)
        self._dones = torch.zeros(n, dtype=torch.bool, device=dev)
        self._terminated = torch.zeros(n, dtype=torch.bool, device=dev)
        self._truncated = torch.zeros(n, dtype=torch.bool, device=dev)
        self._terminal_obs = torch.zeros((n, d), dtype=torch.float32, device=dev) if track_terminal_obs else None
        self._actions: Optional[torch.Tensor] = None
        self._scalars = torch.empty((nat.SC_COUNT, n), dtype=torch.int32, device=dev)
        self._stats = torch.zeros(len(nat.STAT_NAMES), dtype=torch.float64, device=dev)
        self._host: Optional[Dict[str, torch.Tensor]] = None
        # SB3 Monitor's info_keywords (A2C_training.py:124 passes none, so its EvaluationCallback :161-179
        # never sees 'exploration_percentage'): info keys copied into info["episode"] of finished envs
        self.info_keywords = tuple(info_keywords)
        self._last_actions_host: Optional[np.ndarray] = None
        self._t_start = time.time()
        self._waiting = False

        if _spaces is not None:
            self.action_space = _spaces.Discrete(5)                                   # plantos_env.py:41
            self.observation_space = _spaces.Box(low=0, high=1.0, shape=(d,), dtype=np.float32)  # :59-63
        else:
            self.action_space = _Discrete(5)
            self.observation_space = _Box(0, 1.0, (d,), np.float32)

    # ------------------------------------------------------------------ helpers
    def _stream(self) -> int:
        return torch.cuda.current_stream(self.device).cuda_stream

    @property
    def kernel_name(self) -> str:
        return self._lib.plantos_kernel_name(self._h).decode()

    @property
    def last_step_kernel(self) -> str:
        """The kernel the latest step launched (k_step_tile / k_step_fast / k_step_generic)."""
        return self._lib.plantos_last_step_kernel(self._h).decode()

    @property
    def launch_count(self) -> int:
        return int(self._lib.plantos_launch_count(self._h))

    @property
    def state_bytes_per_env(self) -> int:
        return int(self._lib.plantos_state_bytes_per_env(self._h))

    def check(self) -> None:
        """Raise if a device-side error was flagged (synchronises the stream)."""
        nat.check(self._lib.plantos_check(self._h, self._stream()))

    # --------------------------------------------------------------------- maps
    def push_maps(self, cells: np.ndarray, rover: np.ndarray) -> None:
        """Injected-map mode: `cells` u8 [N, E, G, G] cell codes, `rover` [N, E, 2] starts;
        env i consumes map (i, k) at its k-th reset."""
        n, g = self.num_envs, self.grid_size
        cells = np.ascontiguousarray(cells, dtype=np.uint8)
        rover = np.ascontiguousarray(rover, dtype=np.int16)
        if cells.ndim != 4 or cells.shape[0] != n or cells.shape[2:] != (g, g):
            raise ValueError(f"cells must be [N={n}, E, {g}, {g}], got {cells.shape}")
        e = cells.shape[1]
        if rover.shape != (n, e, 2):
            raise ValueError(f"rover must be [{n}, {e}, 2], got {rover.shape}")
        nat.check(self._lib.plantos_push_maps(self._h, cells.ctypes.data, rover.ctypes.data, e))

    # ------------------------------------------------------------- VecEnv surface
    def reset(self) -> torch.Tensor:
        nat.check(self._lib.plantos_reset(self._h, self._obs.data_ptr(), self._stream()))
        self._dones.zero_()
        self._terminated.zero_()
        self._truncated.zero_()
        self._waiting = False
        return self._obs

    def step_async(self, actions) -> None:
        if not isinstance(actions, torch.Tensor):
            actions = torch.as_tensor(np.asarray(actions), dtype=torch.int64)
        if actions.dtype != torch.int64 or actions.device != self.device or not actions.is_contiguous():
            actions = actions.to(device=self.device, dtype=torch.int64, non_blocking=True).contiguous()
        if actions.numel() != self.num_envs:
            raise ValueError(f"expected {self.num_envs} actions, got {actions.numel()}")
        self._actions = actions  # keep alive until the launch has consumed it
        if len(self._obs_ring) > 1:
            self._ring_pos = (self._ring_pos + 1) % len(self._obs_ring)
            self._obs = self._obs_ring[self._ring_pos]
        tobs = self._terminal_obs.data_ptr() if self._terminal_obs is not None else None
        nat.check(self._lib.plantos_step(
            self._h, actions.data_ptr(), self._obs.data_ptr(), self._rewards.data_ptr(),
            self._dones.data_ptr(), self._terminated.data_ptr(), self._truncated.data_ptr(),
            tobs, self._stream()))
        self._waiting = True

    def step_wait(self):
        if not self._waiting:
            raise RuntimeError("step_wait() without step_async()")
        self._waiting = False
        infos = LazyInfos(self)
        if self.full_infos:
            infos = [infos[i] for i in range(self.num_envs)]
        return self._obs, self._rewards, self._dones, infos

    def step(self, actions):
        self.step_async(actions)
        return self.step_wait()

    # ------------------------------------------------------------------ step_many
    def make_rollout(self, k: int, with_flags: bool = False, pipelined: bool = True) -> "GraphRollout":
        """K open-loop steps as ONE CUDA-graph launch: `rollout(actions[K, N])` returns the K
        observation / reward / done tensors (static buffers, overwritten by the next call).  For
        MCTS-style rollouts (mcts_custom_trainer.py:168-243 steps a copied env with a fixed action
        sequence) and for small batches, where the per-step host launch cost dominates.
        `pipelined`: consecutive steps of the rollout overlap (plantos_set_pipelining: the actions of
        all K steps are resident before the launch, every step writes its own observation buffer)."""
        return GraphRollout(self, k, with_flags, pipelined)

    def step_many(self, actions, with_flags: bool = False):
        """K open-loop steps in ONE call (plantos_rollout): `actions` int64 [K, N] -> obs [K, N, D],
        rewards [K, N], dones [K, N] (+ terminated, truncated [K, N] with `with_flags`); the tensors
        are reused by the next call with the same K (one buffer set per K is kept).  Bit-identical to K `step` calls, auto-resets
        included.  On the fast presets the K steps are one launch of the state-resident kernel (the
        envs' window rings and records never leave the SM between steps) -- the GPU form of the
        reference's MCTS rollout loop (mcts_custom_trainer.py:139-166)."""
        if not (isinstance(actions, torch.Tensor) and actions.dtype == torch.int64 and actions.device == self.device
                and actions.is_contiguous()):              # (the fast path costs no torch calls at all)
            if not isinstance(actions, torch.Tensor):
                actions = torch.as_tensor(np.asarray(actions), dtype=torch.int64)
            actions = actions.to(device=self.device, dtype=torch.int64, non_blocking=True).contiguous()
        if actions.dim() != 2 or actions.shape[1] != self.num_envs:
            raise ValueError(f"expected actions of shape [K, {self.num_envs}], got {tuple(actions.shape)}")
        k, n, d, dev = int(actions.shape[0]), self.num_envs, self.obs_dim, self.device
        if not hasattr(self, "_many"):
            self._many = {}                                # one buffer set per rollout length
        buf = self._many.get(k)
        if buf is None:
            stride = (n * d + 3) // 4 * 4                  # per-step stride padded to 16 bytes
            buf = {"k": k, "stride": stride,
                   "obs": torch.empty(k * stride, dtype=torch.float32, device=dev),
                   "rew": torch.empty((k, n), dtype=torch.float32, device=dev),
                   "done": torch.zeros((k, n), dtype=torch.bool, device=dev),
                   "term": torch.zeros((k, n), dtype=torch.bool, device=dev),
                   "trunc": torch.zeros((k, n), dtype=torch.bool, device=dev)}
            buf["obs_view"] = buf["obs"].as_strided((k, n, d), (stride, d, 1))
            buf["ptrs"] = (buf["obs"].data_ptr(), buf["rew"].data_ptr(), buf["done"].data_ptr(),
                           buf["term"].data_ptr(), buf["trunc"].data_ptr(),
                           self._terminal_obs.data_ptr() if self._terminal_obs is not None else None)
            self._many[k] = buf
        po, pr, pd, pt, pu, ptobs = buf["ptrs"]
        rc = self._lib.plantos_rollout(self._h, k, actions.data_ptr(), po, buf["stride"], pr, pd,
                                       pt if with_flags else None, pu if with_flags else None, ptobs, self._stream())
        if rc:
            nat.check(rc)
        self._actions = actions                            # keep alive until the launch has consumed it
        if with_flags:
            return buf["obs_view"], buf["rew"], buf["done"], buf["term"], buf["trunc"]
        return buf["obs_view"], buf["rew"], buf["done"]

    def set_pipelining(self, enable: bool) -> None:
        """Let back-to-back `step_async` calls overlap on the device (open-loop stepping only: the
        actions of a step must not be computed from the previous step's results, and `obs_ring >= 2`
        so that consecutive steps write different observation buffers).  See plantos_set_pipelining."""
        nat.check(self._lib.plantos_set_pipelining(self._h, int(bool(enable))))
        self._pipelining = bool(enable)

    def rollout_policy(self, uniforms: Optional[torch.Tensor] = None, generator: Optional[torch.Generator] = None) -> torch.Tensor:
        """Actions of the reference's MCTS rollout policy (mcts_custom_trainer.py:168-216) for the
        current state of every env: 70 % least-visited valid neighbour, else random.  `uniforms`
        float32 [N, 2] in [0, 1) (drawn with `generator` if omitted); returns int64 [N] on the device."""
        if uniforms is None:
            uniforms = torch.rand((self.num_envs, 2), dtype=torch.float32, device=self.device, generator=generator)
        uniforms = uniforms.to(device=self.device, dtype=torch.float32).contiguous()
        actions = torch.empty(self.num_envs, dtype=torch.int64, device=self.device)
        nat.check(self._lib.plantos_rollout_policy(self._h, uniforms.data_ptr(), actions.data_ptr(), self._stream()))
        self._policy_uniforms = uniforms  # keep alive until the launch has consumed it
        return actions

    def close(self) -> None:
        if getattr(self, "_h", None) is not None and self._h:
            self._lib.plantos_destroy(self._h)
            self._h = None

    def __del__(self):  # pragma: no cover
        try:
            self.close()
        except Exception:
            pass

    def seed(self, seed: Optional[int] = None) -> List[Optional[int]]:
        # PlantOSEnv.reset(seed=) never reaches the map generator (plantos_env.py:127 vs :344);
        # maps here come from the Philox key given at construction.
        return [None] * self.num_envs

    def get_attr(self, attr_name: str, indices=None) -> List[Any]:
        return [getattr(self, attr_name)] * len(self._indices(indices))

    def set_attr(self, attr_name: str, value: Any, indices=None) -> None:
        """The attributes a caller of the reference may change after construction: `max_steps`
        (plantos_env.py:120) and the reward constants `R_*` (:65-93, also by their config names
        `r_*`).  They apply to ALL envs of the batch (`indices` must cover every env) and to every
        step enqueued afterwards.  Everything else is baked into the device state."""
        if indices is not None and sorted(self._indices(indices)) != list(range(self.num_envs)):
            raise ValueError("the batched simulator sets attributes for all envs at once (indices=None)")
        name = attr_name.lower()
        if name == "max_steps":
            nat.check(self._lib.plantos_set_max_steps(self._h, int(value)))
            self.max_steps = int(value)
        elif name in self.rewards:
            self.rewards[name] = float(value)
            rw = tables.reward_table(self.rewards)
            nat.check(self._lib.plantos_upload_tables(self._h, None, None, None, None, rw.ctypes.data))
        else:
            raise AttributeError(f"{attr_name!r} cannot be changed after construction "
                                 f"(settable: max_steps, {', '.join(k.upper() for k in self.rewards)})")

    def env_method(self, method_name: str, *args, indices=None, **kwargs) -> List[Any]:
        """Per-env method calls have no counterpart on the batch; the only methods the reference's
        call sites use through a VecEnv are answered for all envs at once."""
        if method_name in ("get_wrapper_attr", "get_attr"):
            return self.get_attr(args[0], indices)
        if method_name == "close":
            return [None] * len(self._indices(indices))
        raise AttributeError(f"env_method({method_name!r}): PlantOSEnv exposes no such method through the batched simulator")

    def env_is_wrapped(self, wrapper_class, indices=None) -> List[bool]:
        return [False] * len(self._indices(indices))

    def _indices(self, indices) -> List[int]:
        if indices is None:
            return list(range(self.num_envs))
        if isinstance(indices, int):
            return [indices]
        return list(indices)

    # -------------------------------------------------- extra tensors of the last step
    @property
    def terminated(self) -> torch.Tensor:
        return self._terminated

    @property
    def truncated(self) -> torch.Tensor:
        return self._truncated

    @property
    def terminal_observation(self) -> Optional[torch.Tensor]:
        """[N, D]; row i is valid where the last `dones[i]` was True."""
        return self._terminal_obs

    # --------------------------------------------------------------- host-buffer step
    def step_host(self, actions: np.ndarray):
        """numpy in / numpy out, like an SB3 numpy VecEnv: H2D actions, step, D2H results.
        Uses pinned host buffers owned by this object; returns views of them."""
        if self._host is None:
            n, d = self.num_envs, self.obs_dim
            self._host = {
                "actions": torch.empty(n, dtype=torch.int64).pin_memory(),
                "obs": torch.empty((n, d), dtype=torch.float32).pin_memory(),
                "rewards": torch.empty(n, dtype=torch.float32).pin_memory(),
                "dones": torch.empty(n, dtype=torch.bool).pin_memory(),
            }
        hb = self._host
        hb["actions"].numpy()[:] = np.asarray(actions, dtype=np.int64).reshape(-1)
        nat.check(self._lib.plantos_step_host(
            self._h, hb["actions"].data_ptr(), hb["obs"].data_ptr(), hb["rewards"].data_ptr(),
            hb["dones"].data_ptr(), self._stream()))
        return hb["obs"].numpy(), hb["rewards"].numpy(), hb["dones"].numpy()

    # ------------------------------------------------------------------- state / info
    def scalars(self, terminal: bool = False) -> Dict[str, torch.Tensor]:
        """Per-env integer state as int32 tensors (names: _native.SC_NAMES)."""
        nat.check(self._lib.plantos_get_scalars(self._h, int(terminal), self._scalars.data_ptr(), self._stream()))
        out = self._scalars.clone()
        return {name: out[k] for k, name in enumerate(nat.SC_NAMES)}

    def returns(self, terminal: bool = False) -> torch.Tensor:
        out = torch.empty(self.num_envs, dtype=torch.float64, device=self.device)
        nat.check(self._lib.plantos_get_returns(self._h, int(terminal), out.data_ptr(), self._stream()))
        return out

    def get_state(self) -> Dict[str, torch.Tensor]:
        """cells u8 [N,G,G], visits i32 [N,G,G] and the scalar dict."""
        n, g = self.num_envs, self.grid_size
        cells = torch.empty((n, g, g), dtype=torch.uint8, device=self.device)
        visits = torch.empty((n, g, g), dtype=torch.int32, device=self.device)
        nat.check(self._lib.plantos_get_state(self._h, cells.data_ptr(), visits.data_ptr(), self._stream()))
        state = {"cells": cells, "visits": visits}
        state.update(self.scalars())
        return state

    def set_state(self, cells: Optional[torch.Tensor] = None, visits: Optional[torch.Tensor] = None,
                  scalars: Optional[Dict[str, torch.Tensor]] = None) -> None:
        def dev(t, dtype):
            return None if t is None else torch.as_tensor(t).to(device=self.device, dtype=dtype).contiguous()
        cells_t, visits_t = dev(cells, torch.uint8), dev(visits, torch.int32)
        sc_t = None
        if scalars is not None:
            cur = self.scalars()
            cur.update({k: torch.as_tensor(v).to(self.device, torch.int32) for k, v in scalars.items()})
            sc_t = torch.stack([cur[name] for name in nat.SC_NAMES]).contiguous()
        nat.check(self._lib.plantos_set_state(
            self._h, cells_t.data_ptr() if cells_t is not None else None,
            visits_t.data_ptr() if visits_t is not None else None,
            sc_t.data_ptr() if sc_t is not None else None, self._stream()))
        torch.cuda.current_stream(self.device).synchronize()  # inputs may be temporaries

    def _info_snapshot(self) -> Dict[str, Any]:
        live = {k: v.cpu().numpy() for k, v in self.scalars(False).items()}
        done = self._dones.cpu().numpy()
        snap: Dict[str, Any] = {"live": live, "done": done,
                                "terminated": self._terminated.cpu().numpy(),
                                "truncated": self._truncated.cpu().numpy(),
                                "actions": self._actions.cpu().numpy() if self._actions is not None and self._actions.dim() == 1 else None}
        if done.any():
            snap["term"] = {k: v.cpu().numpy() for k, v in self.scalars(True).items()}
            snap["term_return"] = self.returns(True).cpu().numpy()
        return snap

    # --------------------------------------------------------------- episode statistics
    def episode_stats(self, clear: bool = False, all_reduce: bool = True) -> Dict[str, float]:
        """Sums over finished episodes since construction / the last clear.  With
        torch.distributed initialised the 8-vector is summed over ranks (NCCL on the GPU
        ranks) -- the only inter-GPU traffic of the simulator."""
        vals = self.episode_stats_tensor(clear=clear, all_reduce=all_reduce).cpu().tolist()
        return dict(zip(nat.STAT_NAMES, vals))

    def episode_stats_tensor(self, clear: bool = False, all_reduce: bool = True, async_op: bool = False):
        """Same 8-vector as a float64 CUDA tensor, enqueued without a host sync (the kernel
        and the collective are stream-ordered), for callers that poll it off the step path.

        `async_op=True` returns a `PendingStats`: the snapshot is taken on the env's stream now, the
        all-reduce runs on the process group's own stream NEXT TO the steps enqueued afterwards (it is
        8 doubles: pure latency, which the following launches hide); `.wait()` makes the env's stream
        wait for it and returns the tensor."""
        import torch.distributed as dist
        reduce = all_reduce and dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1
        # (with a process group the snapshot goes straight into a fresh 64-byte tensor that is all-reduced in place:
        # no copy kernel between the snapshot and the collective)
        out = torch.empty_like(self._stats) if reduce else self._stats
        nat.check(self._lib.plantos_stats(self._h, out.data_ptr(), int(clear), self._stream()))
        if not reduce:
            return PendingStats(out, None) if async_op else out
        work = dist.all_reduce(out, op=dist.ReduceOp.SUM, async_op=async_op)
        return PendingStats(out, work) if async_op else out

    # --------------------------------------------------------------- per-episode log (SB3 Monitor)
    EPISODE_DTYPE = np.dtype([("env", "<u4"), ("l", "<u4"), ("step_seq", "<u4"), ("flags", "<u4"),
                              ("r", "<f8"), ("collisions", "<u2"), ("watered", "<u2"),
                              ("explored_cells", "<u2"), ("total_cells", "<u2")])

    def enable_episode_log(self, capacity: Optional[int] = None) -> None:
        """Record (env, r, l, ...) of every finished episode on the device -- what SB3's Monitor
        wrapper records per env (A2C_training.py:124).  `capacity` = entries kept between two
        drains (default: 4 per env)."""
        self._ep_cap = int(capacity) if capacity is not None else 4 * self.num_envs
        nat.check(self._lib.plantos_episode_log_enable(self._h, self._ep_cap))
        self._ep_buf = np.zeros(self._ep_cap, dtype=self.EPISODE_DTYPE)

    def curriculum_thresholds(self) -> torch.Tensor:
        """Current exploration threshold of every env (float64 CUDA tensor)."""
        out = torch.empty(self.num_envs, dtype=torch.float64, device=self.device)
        nat.check(self._lib.plantos_get_curriculum_thresholds(self._h, out.data_ptr(), self._stream()))
        return out

    def drain_episode_log(self):
        """Finished episodes since the last drain as a structured array (fields env [global id], l,
        step_seq, flags, r, collisions, watered, explored_cells, total_cells) plus the number of
        entries lost to overflow."""
        n, dropped = C.c_int(0), C.c_int64(0)
        nat.check(self._lib.plantos_episode_log_drain(
            self._h, self._ep_buf.ctypes.data_as(C.c_void_p), self._ep_cap, C.byref(n), C.byref(dropped),
            self._stream()))
        raw = self._ep_buf[:n.value]
        out = np.zeros(n.value, dtype=[("env", "<i8")] + self.EPISODE_DTYPE.descr[1:])
        for name in self.EPISODE_DTYPE.names[1:]:
            out[name] = raw[name]
        out["env"] = raw["env"].astype(np.int64) + self.env_id_base
        return out, int(dropped.value)


class GraphRollout:
    """See PlantOSVecEnv.make_rollout.  The K plantos_step launches are captured once into a
    torch.cuda.CUDAGraph with static action / output buffers; a call copies the actions in and
    replays the graph.  Auto-reset applies inside the rollout exactly as in single steps; infos are
    not produced (read `env.scalars()` afterwards if needed); the device episode log would stamp
    every replay with the step numbers of the capture."""

    def __init__(self, env: "PlantOSVecEnv", k: int, with_flags: bool = False, pipelined: bool = True):
        """`with_flags`: also write the env's terminated / truncated / terminal-observation buffers
        every step, like a single step does (they hold the last step's values afterwards)."""
        if k < 1:
            raise ValueError("k must be >= 1")
        n, d, dev = env.num_envs, env.obs_dim, env.device
        self.env, self.k = env, int(k)
        self.actions = torch.zeros((k, n), dtype=torch.int64, device=dev)
        # [K, N, D] view whose per-step stride is padded to 16 bytes (the fast kernel's row stores)
        stride = (n * d + 3) // 4 * 4
        self.obs = torch.empty(k * stride, dtype=torch.float32, device=dev).as_strided((k, n, d), (stride, d, 1))
        self.rewards = torch.empty((k, n), dtype=torch.float32, device=dev)
        self.dones = torch.zeros((k, n), dtype=torch.bool, device=dev)
        self.graph = torch.cuda.CUDAGraph()
        lib, h = env._lib, env._h
        torch.cuda.synchronize(dev)
        was = getattr(env, "_pipelining", False)
        nat.check(lib.plantos_set_pipelining(h, int(bool(pipelined))))   # baked into the captured launches
        try:
            self._capture(env, with_flags)
        finally:
            nat.check(lib.plantos_set_pipelining(h, int(was)))

    def _capture(self, env: "PlantOSVecEnv", with_flags: bool) -> None:
        lib, h, dev = env._lib, env._h, env.device
        with torch.cuda.graph(self.graph):
            stream = C.c_void_p(torch.cuda.current_stream(dev).cuda_stream)
            term = env._terminated.data_ptr() if with_flags else None
            trunc = env._truncated.data_ptr() if with_flags else None
            tobs = env._terminal_obs.data_ptr() if (with_flags and env._terminal_obs is not None) else None
            for t in range(self.k):
                nat.check(lib.plantos_step(h, self.actions[t].data_ptr(), self.obs[t].data_ptr(),
                                           self.rewards[t].data_ptr(), self.dones[t].data_ptr(), term, trunc, tobs, stream))

    def __call__(self, actions):
        if not isinstance(actions, torch.Tensor):
            actions = torch.as_tensor(np.asarray(actions), dtype=torch.int64)
        if tuple(actions.shape) != tuple(self.actions.shape):
            raise ValueError(f"expected actions of shape {tuple(self.actions.shape)}, got {tuple(actions.shape)}")
        self.actions.copy_(actions, non_blocking=True)
        self.graph.replay()
        return self.obs, self.rewards, self.dones


class MonitorCSV:
    """`monitor.csv` in the format of stable_baselines3.common.monitor.Monitor (header
    `#{"t_start": ..., "env_id": ...}`, then `r,l,t` rows; reference artefacts
    train_improved1/gym/env_*.monitor.csv), fed from the device episode log.

    One file for the whole VecEnv (rows in the order episodes were logged, with an extra leading
    `env` column when `env_column=True`), or one `env_<id>.monitor.csv` per env like the
    reference's `Monitor(env, f"env_{rank}")` when `per_env_files=True` (small N only).
    `t` is the wall-clock time of the `flush()` that collected the episode, relative to t_start:
    steps are enqueued asynchronously, so the host never sees the exact moment an episode ended."""

    # what `info_keywords` may name: info keys of PlantOSEnv._get_info (plantos_env.py:323-336) that the
    # device log carries for the final step of an episode
    INFO_KEYWORDS = {
        "exploration_percentage": lambda ep: (int(ep["explored_cells"]) / int(ep["total_cells"])) * 100,
        "explored_cells": lambda ep: int(ep["explored_cells"]),
        "total_cells": lambda ep: int(ep["total_cells"]),
        "total_collisions": lambda ep: int(ep["collisions"]),
        "plants_watered": lambda ep: int(ep["watered"]),
    }

    def __init__(self, env: "PlantOSVecEnv", directory: str, per_env_files: bool = False,
                 env_column: bool = False, capacity: Optional[int] = None, info_keywords: Sequence[str] = ()):
        import json
        import os
        self.env, self.dir = env, directory
        self.per_env, self.env_column = per_env_files, env_column
        unknown = [k for k in info_keywords if k not in self.INFO_KEYWORDS]
        if unknown:
            raise ValueError(f"info_keywords {unknown} are not recorded per episode (available: {sorted(self.INFO_KEYWORDS)})")
        self.info_keywords = tuple(info_keywords)          # extra columns, like SB3 Monitor(info_keywords=...)
        self.t_start = time.time()
        self.dropped = 0
        os.makedirs(directory, exist_ok=True)
        env.enable_episode_log(capacity)
        self._files: Dict[Any, Any] = {}
        self._json, self._os = json, os

    def _file(self, key):
        f = self._files.get(key)
        if f is None:
            name = "monitor.csv" if key is None else f"env_{key}.monitor.csv"
            f = open(self._os.path.join(self.dir, name), "w")
            f.write("#%s\n" % self._json.dumps({"t_start": self.t_start, "env_id": "None" if key is None else str(key)}))
            f.write(("env," if (key is None and self.env_column) else "") + ",".join(("r", "l", "t") + self.info_keywords) + "\n")
            self._files[key] = f
        return f

    def flush(self) -> int:
        """Drain the device log and append its episodes; returns how many were written."""
        eps, dropped = self.env.drain_episode_log()
        self.dropped += dropped
        t = round(time.time() - self.t_start, 6)
        for ep in eps:
            f = self._file(int(ep["env"]) if self.per_env else None)
            lead = f"{int(ep['env'])}," if (not self.per_env and self.env_column) else ""
            extra = "".join("," + str(self.INFO_KEYWORDS[k](ep)) for k in self.info_keywords)
            f.write(f"{lead}{round(float(ep['r']), 6)},{int(ep['l'])},{t}{extra}\n")
        for f in self._files.values():
            f.flush()
        return len(eps)

    def close(self) -> None:
        self.flush()
        for f in self._files.values():
            f.close()
        self._files.clear()


class PendingStats:
    """A statistics vector whose all-reduce may still be in flight (`all_reduce_stats(async_op=True)`)."""

    def __init__(self, tensor: torch.Tensor, work):
        self.tensor, self._work = tensor, work

    def wait(self) -> torch.Tensor:
        """Order the current stream (NCCL) / the caller (gloo) after the collective; returns the summed vector."""
        if self._work is not None:
            self._work.wait()
            self._work = None
        return self.tensor


def all_reduce_stats(vec: torch.Tensor, async_op: bool = False):
    """Sum a rank-local statistics vector over the default process group (no-op if there is
    none).  Works on NCCL (CUDA tensor) and gloo (CPU tensor) alike.  `async_op=True` returns a
    `PendingStats` instead of the tensor."""
    import torch.distributed as dist
    work = None
    if dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1:
        vec = vec.clone()
        work = dist.all_reduce(vec, op=dist.ReduceOp.SUM, async_op=async_op)
    return PendingStats(vec, work) if async_op else vec


def shard_range(total_envs: int, rank: int, world_size: int):
    """Contiguous block of global env ids owned by `rank`: [start, start + count)."""
    if not 0 <= rank < world_size:
        raise ValueError("rank out of range")
    base, rem = divmod(int(total_envs), int(world_size))
    count = base + (1 if rank < rem else 0)
    start = rank * base + min(rank, rem)
    return start, count
