"""GPU parity proper: the CUDA path (through the C ABI) replays the trajectories recorded
from the unmodified reference.  Integer state, flags and LIDAR hits must be bit-exact;
observations and rewards within 1e-5 absolute (BASELINE.json north_star) -- and the harness
also reports whether they were in fact bit-equal."""
import os

import numpy as np
import pytest

from replay import FIXTURES, check_replay, load_fixture

pytestmark = pytest.mark.gpu

# fixtures whose (G, R, C) has a fast-kernel instantiation
FAST_FIXTURES = ["replay_T_8env", "replay_DFLT_1env"]


def _run(name, kernel, replicas=1, **kw):
    from gpu_backend import GpuBackend
    fx = load_fixture(name)
    be = GpuBackend(fx, kernel=kernel, replicas=replicas)
    try:
        assert be.env.kernel_name == kernel
        res = check_replay(fx, be, **kw)
    finally:
        be.close()
    assert res["episodes"] == len(fx["term_t"]) or kw.get("steps")
    return res


@pytest.mark.parametrize("name", FIXTURES)
def test_generic_kernel_replays_reference(name):
    res = _run(name, "generic")
    assert res["bitexact_obs"] == 1 and res["bitexact_reward"] == 1


@pytest.mark.parametrize("name", FAST_FIXTURES)
def test_fast_kernel_replays_reference(name):
    # fewer than 32 envs: the fast kernel's ragged-tail path
    res = _run(name, "fast")
    assert res["bitexact_obs"] == 1 and res["bitexact_reward"] == 1


@pytest.mark.parametrize("name,replicas", [("replay_T_8env", 13), ("replay_DFLT_1env", 70)])
def test_fast_kernel_pipeline_replays_reference(name, replicas):
    """The tiled pipeline proper: the fixture's envs replicated to fill many warp tiles
    (104 envs = 26 tiles of 4; 70 = 17 tiles + 2 tail envs), every copy checked."""
    res = _run(name, "fast", replicas=replicas)
    assert res["bitexact_obs"] == 1 and res["bitexact_reward"] == 1


@pytest.mark.parametrize("name,kernel,replicas", [("replay_curr_a2c_4env", "fast", 23), ("replay_curr_a2c_4env", "generic", 1),
                                                  ("replay_curr_dqn_3env", "generic", 1)])
def test_device_curriculum_replays_reference_wrapper(name, kernel, replicas):
    """CurriculumWrapper on the device (plantos_set_curriculum; both step kernels) against trajectories
    recorded through the reference's own wrapper class (A2C_training.py:37-109 / trainingCode.py:24-98):
    visit counts that persist across episodes, threshold terminations, reset observations that show
    fresh counts; plus the per-env thresholds at the end."""
    import torch
    from gpu_backend import GpuBackend
    fx = load_fixture(name)
    be = GpuBackend(fx, kernel=kernel, replicas=replicas)
    assert be.env.kernel_name == kernel
    res = check_replay(fx, be)
    assert res["bitexact_obs"] == 1 and res["bitexact_reward"] == 1 and res["episodes"] >= 40
    # after the last step (and its auto-resets) the thresholds equal the wrapper's
    want = fx["cur_threshold"][-1].copy()
    done_last = fx["terminated"][-1] | fx["truncated"][-1]
    got = be.env.curriculum_thresholds().cpu().numpy()[:len(want)]
    assert np.array_equal(got[~done_last], want[~done_last])
    with pytest.raises(Exception):
        be.env.set_state(cells=torch.zeros(1))                 # refused while a curriculum is active


@pytest.mark.parametrize("impl", ["tile", "trip"])
@pytest.mark.parametrize("grid,replicas", [(1, 60), (2, 60), (1, 24), (5, 100), (1, 200), (2, 131)])
def test_fast_kernel_buffer_reuse(grid, replicas, impl, monkeypatch):
    """Few persistent blocks => every warp walks several 32-env macro tiles (480 envs over 7
    warps = tiles of 32, 32, 8; over 14 warps = 32, 4), so the record / target-word buffers and
    both window buffers are reused and the next-tile prefetch path runs; (1, 24) and (5, 100)
    are single short macro tiles of 28 and 24 envs."""
    # k_step_tile (16-warp blocks, 32-env tiles): (1, 200) = 50 tiles over 16 warps, three to four per warp
    # (both record buffers and the window buffer are reused), (2, 131) ends in a short tile of 24 envs
    monkeypatch.setenv("PLANTOS_FAST_GRID", str(grid))
    monkeypatch.setenv("PLANTOS_FAST_IMPL", impl)
    res = _run("replay_T_8env", "fast", replicas=replicas, steps=1100)
    assert res["bitexact_obs"] == 1


def test_c_abi_direct_step_host():
    """The boundary a non-Python host binds: raw ctypes calls with HOST buffers
    (plantos_step_host), no VecEnv object in between."""
    import ctypes as C
    import torch
    from rl_env_b200 import _native as nat
    fx = load_fixture("replay_DFLT_1env")
    lib = nat.load()
    cfg = nat.Config()
    nat.check(lib.plantos_default_config(C.byref(cfg)))  # ctor defaults == this fixture's config
    cfg.num_envs = 1
    cfg.map_source = nat.MAPS_INJECTED
    h = C.c_void_p()
    nat.check(lib.plantos_create(C.byref(cfg), 0, C.byref(h)))
    try:
        d = lib.plantos_obs_dim(C.byref(cfg))
        assert d == 77
        cells = np.ascontiguousarray(fx["maps_cells"])
        rover = np.ascontiguousarray(fx["maps_rover"])
        nat.check(lib.plantos_push_maps(h, cells.ctypes.data, rover.ctypes.data, cells.shape[1]))
        obs_dev = torch.empty((1, d), dtype=torch.float32, device="cuda:0")
        nat.check(lib.plantos_reset(h, obs_dev.data_ptr(), None))
        torch.cuda.synchronize()
        assert np.array_equal(obs_dev.cpu().numpy(), fx["reset_obs"])
        obs = np.zeros((1, d), np.float32)
        rew = np.zeros(1, np.float32)
        done = np.zeros(1, np.uint8)
        for t in range(1200):
            act = np.ascontiguousarray(fx["actions"][t])
            nat.check(lib.plantos_step_host(h, act.ctypes.data, obs.ctypes.data, rew.ctypes.data,
                                            done.ctypes.data, None))
            assert np.array_equal(obs, fx["obs"][t]), t
            assert rew[0] == np.float32(fx["rewards"][t, 0]), t
            assert bool(done[0]) == bool(fx["terminated"][t, 0] or fx["truncated"][t, 0]), t
        nat.check(lib.plantos_check(h, None))
        assert lib.plantos_launch_count(h) >= 1201
    finally:
        lib.plantos_destroy(h)


def test_out_of_maps_is_reported():
    from rl_env_b200 import PlantOSError, PlantOSVecEnv
    fx = load_fixture("replay_odd_4env")
    from replay import fixture_kwargs
    env = PlantOSVecEnv(4, map_source="injected", max_steps=5, **fixture_kwargs(fx))
    env.push_maps(fx["maps_cells"][:, :2], fx["maps_rover"][:, :2])
    env.reset()
    for t in range(12):  # 2 auto-resets per env > 1 spare map
        env.step(fx["actions"][t])
    with pytest.raises(PlantOSError) as ei:
        env.check()
    assert ei.value.code == -3
    env.close()


def test_no_cpu_fallback_and_bad_configs():
    from rl_env_b200 import PlantOSError, PlantOSVecEnv
    with pytest.raises(ValueError):
        PlantOSVecEnv(4, device="cpu")
    with pytest.raises(PlantOSError):
        PlantOSVecEnv(4, grid_size=6, num_plants=30)     # the reference's ValueError (plantos_env.py:360)
    with pytest.raises(PlantOSError):
        PlantOSVecEnv(4, kernel="fast", grid_size=64, lidar_range=32, lidar_channels=16, num_plants=8)
    env = PlantOSVecEnv(4)
    with pytest.raises(PlantOSError):
        env.step(np.zeros(4, np.int64))                  # step before reset
    env.close()
