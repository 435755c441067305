"""ctypes binding of include/plantos.h (libplantos_b200.so).  No CPU fallback: if the
library cannot be loaded the import of the simulator fails loudly."""
from __future__ import annotations

import ctypes as C
import os
from typing import Optional

from . import build as _build

ABI_VERSION = 2

OK, EINVAL, ECUDA, ENOMAPS, ESTATE = 0, -1, -2, -3, -4
MAPS_PHILOX, MAPS_INJECTED, MAPS_MAZE = 0, 1, 2
KERNEL_AUTO, KERNEL_GENERIC, KERNEL_FAST = 0, 1, 2
RW_COUNT = 6
SC_NAMES = ("x", "y", "step_count", "explored_cells", "total_cells", "thirsty_plants",
            "total_collisions", "collided_with_wall", "completion_bonus_given", "episode", "watered")
SC_COUNT = len(SC_NAMES)
STAT_NAMES = ("episodes", "return_sum", "length_sum", "exploration_pct_sum", "collisions_sum",
              "watered_sum", "terminated", "truncated")


CURRICULUM_OFF, CURRICULUM_TERMINATE, CURRICULUM_MARK = 0, 1, 2


class Config(C.Structure):
    """plantos_config_t"""
    _fields_ = [
        ("struct_size", C.c_int32), ("num_envs", C.c_int32), ("env_id_base", C.c_int64),
        ("grid_size", C.c_int32), ("num_plants", C.c_int32), ("num_obstacles", C.c_int32),
        ("lidar_range", C.c_int32), ("lidar_channels", C.c_int32), ("max_steps", C.c_int32),
        ("thirsty_plant_prob", C.c_float), ("map_source", C.c_int32), ("seed", C.c_uint64),
        ("r_goal", C.c_double), ("r_mistake", C.c_double), ("r_invalid", C.c_double),
        ("r_water_empty", C.c_double), ("r_step", C.c_double), ("r_exploration", C.c_double),
        ("r_revisit", C.c_double), ("r_complete_exploration", C.c_double),
        ("kernel", C.c_int32), ("tune_fast_grid", C.c_int32), ("tune_fast_impl", C.c_int32),
        ("tune_no_pdl", C.c_int32), ("tune_l2_keep_mb", C.c_int32), ("reserved", C.c_int32 * 3),
    ]


class PlantOSError(RuntimeError):
    def __init__(self, code: int, message: str):
        super().__init__(f"libplantos_b200 error {code}: {message}")
        self.code = code


_vp = C.c_void_p
# name -> (restype, argtypes); every symbol include/plantos.h declares
SIGNATURES = {
    "plantos_default_config": (C.c_int, [C.POINTER(Config)]),
    "plantos_obs_dim": (C.c_int, [C.POINTER(Config)]),
    "plantos_compute_tables": (C.c_int, [C.POINTER(Config), _vp, _vp, _vp, _vp, _vp]),
    "plantos_create": (C.c_int, [C.POINTER(Config), C.c_int, C.POINTER(_vp)]),
    "plantos_destroy": (C.c_int, [_vp]),
    "plantos_upload_tables": (C.c_int, [_vp, _vp, _vp, _vp, _vp, _vp]),
    "plantos_push_maps": (C.c_int, [_vp, _vp, _vp, C.c_int]),
    "plantos_reset": (C.c_int, [_vp, _vp, _vp]),
    "plantos_step": (C.c_int, [_vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp]),
    "plantos_rollout": (C.c_int, [_vp, C.c_int, _vp, _vp, C.c_int64, _vp, _vp, _vp, _vp, _vp, _vp]),
    "plantos_step_host": (C.c_int, [_vp, _vp, _vp, _vp, _vp, _vp]),
    "plantos_get_scalars": (C.c_int, [_vp, C.c_int, _vp, _vp]),
    "plantos_get_returns": (C.c_int, [_vp, C.c_int, _vp, _vp]),
    "plantos_get_state": (C.c_int, [_vp, _vp, _vp, _vp]),
    "plantos_set_state": (C.c_int, [_vp, _vp, _vp, _vp, _vp]),
    "plantos_stats": (C.c_int, [_vp, _vp, C.c_int, _vp]),
    "plantos_check": (C.c_int, [_vp, _vp]),
    "plantos_set_curriculum": (C.c_int, [_vp, C.c_int, C.c_double, C.c_double, C.c_double, C.c_int]),
    "plantos_get_curriculum_thresholds": (C.c_int, [_vp, _vp, _vp]),
    "plantos_rollout_policy": (C.c_int, [_vp, _vp, _vp, _vp]),
    "plantos_episode_log_enable": (C.c_int, [_vp, C.c_int]),
    "plantos_episode_log_drain": (C.c_int, [_vp, _vp, C.c_int, C.POINTER(C.c_int), C.POINTER(C.c_int64), _vp]),
    "plantos_launch_count": (C.c_int64, [_vp]),
    "plantos_kernel_name": (C.c_char_p, [_vp]),
    "plantos_last_step_kernel": (C.c_char_p, [_vp]),
    "plantos_set_pipelining": (C.c_int, [_vp, C.c_int]),
    "plantos_set_curriculum_reuse_map": (C.c_int, [_vp, C.c_int]),
    "plantos_set_max_steps": (C.c_int, [_vp, C.c_int]),
    "plantos_state_bytes_per_env": (C.c_int64, [_vp]),
    "plantos_last_error": (C.c_char_p, []),
    "plantos_abi_version": (C.c_int, []),
}

_lib: Optional[C.CDLL] = None


def lib_path() -> str:
    return _build.LIB_PATH


def load(build_if_missing: bool = True) -> C.CDLL:
    """dlopen the in-tree library (building it first only if it is absent)."""
    global _lib
    if _lib is not None:
        return _lib
    path = lib_path()
    if not os.path.exists(path):
        if not build_if_missing:
            raise ImportError(f"{path} is missing; run `python -m rl_env_b200.build`")
        _build.build(force=True)
    lib = C.CDLL(path)
    for name, (restype, argtypes) in SIGNATURES.items():
        fn = getattr(lib, name)  # AttributeError if the symbol is not exported
        fn.restype = restype
        fn.argtypes = argtypes
    got = lib.plantos_abi_version()
    if got != ABI_VERSION:
        raise ImportError(f"{path}: ABI version {got}, binding expects {ABI_VERSION}")
    _lib = lib
    return lib


def check(code: int) -> None:
    if code != OK:
        msg = load().plantos_last_error()
        raise PlantOSError(code, msg.decode() if msg else "")
