"""CPU-only checks of the C-ABI shared library: it loads, exports every symbol that
include/plantos.h declares, and its host-only entry points (config, tables, validation)
agree with the Python mirror and with both oracles.  No GPU compute is called here."""
import ctypes as C
import os
import re

import numpy as np
import pytest

from rl_env_b200 import _native as nat
from rl_env_b200 import tables

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HEADER = os.path.join(ROOT, "include", "plantos.h")
PRESETS = [(25, 10, 12, 6, 16), (21, 8, 50, 2, 10), (64, 64, 600, 32, 16), (7, 3, 3, 8, 8), (12, 5, 9, 4, 5)]


def _cfg(g, p, o, r, c):
    lib = nat.load()
    cfg = nat.Config()
    assert lib.plantos_default_config(C.byref(cfg)) == 0
    cfg.grid_size, cfg.num_plants, cfg.num_obstacles, cfg.lidar_range, cfg.lidar_channels = g, p, o, r, c
    return lib, cfg


def test_library_exports_every_declared_symbol():
    lib = nat.load()
    text = open(HEADER).read()
    declared = set(re.findall(r"\b(plantos_[a-z_]+)\s*\(", text))
    declared -= {"plantos_config_t"}
    assert declared == set(nat.SIGNATURES), declared ^ set(nat.SIGNATURES)
    for name in declared:
        assert hasattr(lib, name), name
    assert lib.plantos_abi_version() == nat.ABI_VERSION
    # struct layout seen by ctypes == the library's
    cfg = nat.Config()
    assert lib.plantos_default_config(C.byref(cfg)) == 0
    assert cfg.struct_size == C.sizeof(nat.Config)
    # enum values the binding hard-codes
    for name, val in (("PLANTOS_RW_COUNT", nat.RW_COUNT), ("PLANTOS_SC_COUNT", nat.SC_COUNT)):
        assert re.search(rf"{name}\s*=\s*{val}\b", text), name


def test_default_config_is_the_reference_ctor():
    lib, cfg = _cfg(21, 8, 50, 2, 10)
    # plantos_env.py:25-26, :76-83, :120
    assert (cfg.grid_size, cfg.num_plants, cfg.num_obstacles, cfg.lidar_range, cfg.lidar_channels) == (21, 8, 50, 2, 10)
    assert abs(cfg.thirsty_plant_prob - 0.7) < 1e-7 and cfg.max_steps == 1000
    assert (cfg.r_goal, cfg.r_mistake, cfg.r_invalid, cfg.r_water_empty) == (20, -10, -5, -5)
    assert (cfg.r_step, cfg.r_exploration, cfg.r_revisit, cfg.r_complete_exploration) == (-0.1, 10, -1, 50)
    assert lib.plantos_obs_dim(C.byref(cfg)) == 77
    cfg.lidar_channels = 16
    assert lib.plantos_obs_dim(C.byref(cfg)) == 107


@pytest.mark.parametrize("g,p,o,r,c", PRESETS)
def test_tables_agree_across_c_python_and_oracles(g, p, o, r, c):
    from oracle import c_oracle, plantos_oracle
    lib, cfg = _cfg(g, p, o, r, c)
    off = np.zeros((c, r, 2), np.int8)
    dist = np.zeros(r + 1, np.float32)
    pos = np.zeros(g, np.float32)
    vis = np.zeros(11, np.float32)
    rw = np.zeros(12, np.float64)
    assert lib.plantos_compute_tables(C.byref(cfg), off.ctypes.data, dist.ctypes.data, pos.ctypes.data,
                                      vis.ctypes.data, rw.ctypes.data) == 0
    assert np.array_equal(off, tables.lidar_offsets(c, r))
    assert np.array_equal(off, plantos_oracle.lidar_offsets(c, r))
    assert np.array_equal(off, c_oracle.lidar_offsets(c, r))
    assert np.array_equal(dist.view(np.uint32), tables.distance_table(r).view(np.uint32))
    assert np.array_equal(pos.view(np.uint32), tables.position_table(g).view(np.uint32))
    assert np.array_equal(vis.view(np.uint32), tables.visit_table().view(np.uint32))
    assert np.array_equal(rw, tables.reward_table(tables.DEFAULT_REWARDS))
    # device-side alternative (float division) would give the same bits: x/G, r/R, k/10
    assert np.array_equal(pos, (np.arange(g, dtype=np.float32) / np.float32(g)))
    assert np.array_equal(dist, (np.arange(r + 1, dtype=np.float32) / np.float32(r)))


def test_reward_table_bit_patterns():
    rw = tables.reward_table(tables.DEFAULT_REWARDS).astype(np.float32).view(np.uint32)
    # 9.9, -1.1, -5.1, 19.9, -5.1, -10.1 and the +50 twins 59.9, 48.9
    assert [hex(v) for v in rw[:6]] == ["0x411e6666", "0xbf8ccccd", "0xc0a33333", "0x419f3333", "0xc0a33333", "0xc121999a"]
    assert hex(rw[6]) == "0x426f999a" and hex(rw[7]) == "0x4243999a"


def test_training_preset_offsets_known_values():
    off = tables.lidar_offsets(16, 6)
    assert off[0].tolist() == [[1, 0], [2, 0], [3, 0], [4, 0], [5, 0], [6, 0]]
    assert off[1].tolist() == [[0, 0], [1, 0], [2, 1], [3, 1], [4, 1], [5, 2]]
    assert off[2].tolist() == [[0, 0], [1, 1], [2, 2], [2, 2], [3, 3], [4, 4]]
    assert off[4].tolist() == [[0, 1], [0, 2], [0, 3], [0, 4], [0, 5], [0, 6]]
    assert sum(1 for i in range(16) if off[i, 0].tolist() == [0, 0]) == 12   # rover's own cell at r=1


def test_validation_and_loud_failure_without_gpu():
    lib, cfg = _cfg(21, 8, 50, 2, 10)
    h = C.c_void_p()
    bad = [("grid_size", 4), ("grid_size", 129), ("lidar_range", 0), ("lidar_channels", 65), ("max_steps", 70000),
           ("num_envs", 0), ("num_plants", 100), ("map_source", 3), ("kernel", 9), ("struct_size", 8)]
    for field, val in bad:
        _, c2 = _cfg(21, 8, 50, 2, 10)
        setattr(c2, field, val)
        assert lib.plantos_create(C.byref(c2), 0, C.byref(h)) == nat.EINVAL, field
        assert lib.plantos_last_error()
    import torch
    if not torch.cuda.is_available():
        # no device: creation must fail with a CUDA error, never fall back to a CPU path
        assert lib.plantos_create(C.byref(cfg), 0, C.byref(h)) == nat.ECUDA
        from rl_env_b200 import PlantOSVecEnv
        with pytest.raises(RuntimeError):
            PlantOSVecEnv(4)
    with pytest.raises(ValueError):
        from rl_env_b200 import PlantOSVecEnv
        PlantOSVecEnv(4, device="cpu")


def test_product_never_imports_the_oracle():
    pkg = os.path.join(ROOT, "rl_env_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                src = open(os.path.join(dirpath, f)).read()
                # the product package neither imports nor loads anything under oracle/ (a comment may
                # point at the host mirror of the Philox generator)
                assert not re.search(r"^\s*(from|import)\s+oracle", src, re.M), f
                assert "plantos_oracle" not in src and "ref_shim" not in src, f
