"""GPU parity at sizes the golden fixtures do not reach.

* lock-step against the C oracle with thousands of envs and Philox-generated maps (the maps
  the device draws at every reset are read back and injected into the oracle, so the
  comparison covers transition, observation and auto-reset, not the generator);
* the Philox generator itself: bit-identical to its host mirror, invariant to sharding,
  distributionally equal to the reference's `random`-driven generator;
* BASELINE.json's full size (131 072 envs per GPU): size-independent properties and
  generic-vs-fast kernel agreement.
"""
import random

import numpy as np
import pytest

pytestmark = pytest.mark.gpu

T_KW = dict(grid_size=25, num_plants=10, num_obstacles=12, lidar_range=6, lidar_channels=16)
DFLT_KW = dict(grid_size=21, num_plants=8, num_obstacles=50, lidar_range=2, lidar_channels=10)
XL_KW = dict(grid_size=64, num_plants=64, num_obstacles=600, lidar_range=32, lidar_channels=16)


def _lockstep(kw, n, steps, max_steps, kernel, seed, check_planes_every=25):
    import torch
    from oracle.c_oracle import COracle
    from rl_env_b200 import PlantOSVecEnv
    env = PlantOSVecEnv(n, kernel=kernel, seed=seed, max_steps=max_steps, full_infos=False, **kw)
    ora = COracle(n, kw["grid_size"], kw["num_plants"], kw["num_obstacles"], kw["lidar_range"],
                  kw["lidar_channels"], max_steps)
    obs = env.reset().cpu().numpy()
    st = env.get_state()
    cells, xs, ys = st["cells"].cpu().numpy(), st["x"].cpu().numpy(), st["y"].cpu().numpy()
    for i in range(n):
        ora.reset_one(i, cells[i], (xs[i], ys[i]))
    assert np.array_equal(obs.view(np.uint32), ora.obs.view(np.uint32)), "reset observations differ"
    rng = np.random.default_rng(seed)
    n_done = 0
    for t in range(steps):
        actions = rng.integers(0, 5, size=n).astype(np.int64)
        if t % 7 == 3:
            actions[rng.integers(0, n, size=max(1, n // 50))] = 4  # extra watering
        g_obs, g_rew, g_done, _ = env.step(torch.from_numpy(actions).cuda())
        o_obs, o_rew, o_term, o_trunc = ora.step(actions)
        g_term, g_trunc = env.terminated.cpu().numpy(), env.truncated.cpu().numpy()
        assert np.array_equal(g_term, o_term), f"step {t}: terminated differs"
        assert np.array_equal(g_trunc, o_trunc), f"step {t}: truncated differs"
        g_rew = g_rew.cpu().numpy()
        assert np.array_equal(g_rew.view(np.uint32), o_rew.astype(np.float32).view(np.uint32)), f"step {t}: reward"
        done = o_term | o_trunc
        g_obs_np = g_obs.cpu().numpy()
        if done.any():
            n_done += int(done.sum())
            tob = env.terminal_observation.cpu().numpy()
            assert np.array_equal(tob[done].view(np.uint32), ora.terminal_obs[done].view(np.uint32)), \
                f"step {t}: terminal observation"
            tsc = {k: v.cpu().numpy() for k, v in env.scalars(True).items()}
            names = ["x", "y", "step_count", "explored_cells", "total_cells", "thirsty_plants",
                     "total_collisions", "collided_with_wall", "completion_bonus_given"]
            for k, name in enumerate(names):
                assert np.array_equal(tsc[name][done], ora.term_sc[k][done]), f"step {t}: terminal {name}"
            gret = env.returns(True).cpu().numpy()
            assert np.array_equal(gret[done], ora.ep_return[done]), f"step {t}: episode return (f64, exact)"
            st = env.get_state()
            cells, xs, ys = st["cells"].cpu().numpy(), st["x"].cpu().numpy(), st["y"].cpu().numpy()
            for i in np.nonzero(done)[0]:
                ora.reset_one(int(i), cells[i], (xs[i], ys[i]))
        assert np.array_equal(g_obs_np.view(np.uint32), ora.obs.view(np.uint32)), f"step {t}: observation"
        if t % check_planes_every == 0 or t == steps - 1:
            st = env.get_state()
            o_cells, o_visits, o_sc = ora.get_state()
            assert np.array_equal(st["cells"].cpu().numpy(), o_cells), f"step {t}: cell planes"
            assert np.array_equal(st["visits"].cpu().numpy(), o_visits), f"step {t}: visit counts"
            for k, name in enumerate(["x", "y", "step_count", "explored_cells", "total_cells", "thirsty_plants",
                                      "total_collisions", "collided_with_wall", "completion_bonus_given"]):
                assert np.array_equal(st[name].cpu().numpy(), o_sc[k]), f"step {t}: {name}"
    env.check()
    stats = env.episode_stats(all_reduce=False)
    assert stats["episodes"] == n_done
    env.close()
    return n_done


@pytest.mark.parametrize("kernel", ["generic", "fast"])
def test_lockstep_training_preset_4096(kernel):
    assert _lockstep(T_KW, 4096, 150, 60, kernel, seed=3) >= 2 * 4096


@pytest.mark.parametrize("kernel", ["generic", "fast"])
def test_lockstep_default_preset_ragged(kernel):
    # N not a multiple of the warp tile / 4-env store group
    assert _lockstep(DFLT_KW, 1027, 120, 50, kernel, seed=4) >= 2 * 1027


@pytest.mark.parametrize("r,c", [(4, 16), (4, 8)])
def test_lockstep_other_fast_instantiations(r, c):
    # the remaining (R, C) shapes the fast kernel is instantiated for, ragged N, frequent resets
    kw = dict(grid_size=14, num_plants=5, num_obstacles=9, lidar_range=r, lidar_channels=c)
    assert _lockstep(kw, 1543, 140, 35, "fast", seed=8) >= 3 * 1543


def test_lockstep_fast_many_rounds_per_warp(monkeypatch):
    # two persistent blocks => every warp walks ~9 macro tiles of 32 envs (steady-state prefetch of the
    # next tile's records and target words, short last tile), with auto-resets in between
    monkeypatch.setenv("PLANTOS_FAST_GRID", "2")
    assert _lockstep(T_KW, 4000, 130, 45, "fast", seed=9) >= 2 * 4000


def test_lockstep_xl_stress():
    assert _lockstep(XL_KW, 256, 90, 40, "generic", seed=5, check_planes_every=30) >= 2 * 256


def test_lockstep_small_grid_terminations():
    # tiny grid: full exploration, completion bonus, terminated & truncated on the same step
    kw = dict(grid_size=6, num_plants=2, num_obstacles=3, lidar_range=3, lidar_channels=12)
    assert _lockstep(kw, 512, 400, 120, "generic", seed=6) > 512


def test_philox_maps_match_host_mirror_and_sharding():
    from oracle.philox_mapgen import generate_map
    from rl_env_b200 import PlantOSVecEnv
    for kw, base in ((T_KW, 0), (DFLT_KW, (1 << 33) + 5), (XL_KW, 7)):
        env = PlantOSVecEnv(8, seed=0xABCDEF0123, env_id_base=base, max_steps=3, **kw)
        env.reset()
        for episode in range(3):
            st = env.get_state()
            cells, xs, ys = st["cells"].cpu().numpy(), st["x"].cpu().numpy(), st["y"].cpu().numpy()
            for i in range(8):
                want_cells, want_rover = generate_map(0xABCDEF0123, base + i, episode, kw["grid_size"],
                                                      kw["num_plants"], kw["num_obstacles"])
                assert np.array_equal(cells[i], want_cells), (kw, i, episode)
                assert (xs[i], ys[i]) == want_rover
            for _ in range(3):  # max_steps=3 -> auto-reset into the next episode
                env.step(np.full(8, 4, np.int64))
        env.close()
    # shard k of a 16-env job == envs [8, 16) of the unsharded job
    a = PlantOSVecEnv(16, seed=9, **T_KW)
    b = PlantOSVecEnv(8, seed=9, env_id_base=8, **T_KW)
    oa, ob = a.reset().cpu().numpy(), b.reset().cpu().numpy()
    assert np.array_equal(oa[8:], ob)
    assert np.array_equal(a.get_state()["cells"].cpu().numpy()[8:], b.get_state()["cells"].cpu().numpy())
    a.close(); b.close()


def test_device_maze_maps_match_host_mirror_and_lockstep():
    """map_source="maze": the Gradio fork's maze generator (gradio-app/plantos_env_new.py:408-604) on the device.
    (1) The maps are bit-identical to the host mirror (oracle.philox_mapgen.generate_maze_map), episode after
    episode, on the fast and the generic kernel, for the training preset and for a 64x64 grid (10x10 meta grid).
    (2) Lock-step against the C oracle on those maps (read back after every reset) -- long corridors and rooms
    are a different LIDAR workload than the cluster maps."""
    from oracle.c_oracle import COracle
    from oracle.philox_mapgen import generate_maze_map
    from rl_env_b200 import PlantOSVecEnv
    for kw, kernel, n in ((T_KW, "fast", 8), (T_KW, "generic", 8), (XL_KW, "generic", 4)):
        env = PlantOSVecEnv(n, seed=77, env_id_base=3, max_steps=4, map_source="maze", kernel=kernel, **kw)
        env.reset()
        for episode in range(3):
            st = env.get_state()
            cells, xs, ys = st["cells"].cpu().numpy(), st["x"].cpu().numpy(), st["y"].cpu().numpy()
            for i in range(n):
                want_cells, want_rover = generate_maze_map(77, 3 + i, episode, kw["grid_size"], kw["num_plants"], kw["num_obstacles"])
                assert np.array_equal(cells[i], want_cells), (kernel, i, episode)
                assert (xs[i], ys[i]) == want_rover
            for _ in range(4):
                env.step(np.full(n, 4, np.int64))
        env.check(); env.close()
    n, steps, max_steps = 2048, 90, 30
    kw = T_KW
    env = PlantOSVecEnv(n, seed=5, max_steps=max_steps, map_source="maze", kernel="fast", **kw)
    ora = COracle(n, kw["grid_size"], kw["num_plants"], kw["num_obstacles"], kw["lidar_range"], kw["lidar_channels"], max_steps)
    obs = env.reset().cpu().numpy()
    st = env.get_state()
    cells, xs, ys = st["cells"].cpu().numpy(), st["x"].cpu().numpy(), st["y"].cpu().numpy()
    for i in range(n):
        ora.reset_one(i, cells[i], (xs[i], ys[i]))
    assert np.array_equal(obs.view(np.uint32), ora.obs.view(np.uint32))
    free = (cells != 1).reshape(n, -1).sum(1)
    assert 400 < free.mean() < 600                         # 16 rooms + corridors on 625 cells: ~16 % walls in long straight runs
    rng = np.random.default_rng(2)
    for t in range(steps):
        a = rng.integers(0, 5, size=n).astype(np.int64)
        g_obs, g_rew, g_done, _ = env.step(a)
        o_obs, o_rew, o_term, o_trunc = ora.step(a)
        done = o_term | o_trunc
        assert np.array_equal(g_done.cpu().numpy(), done), t
        assert np.array_equal(g_rew.cpu().numpy(), o_rew.astype(np.float32)), t
        if done.any():
            st = env.get_state()
            cells, xs, ys = st["cells"].cpu().numpy(), st["x"].cpu().numpy(), st["y"].cpu().numpy()
            for i in np.nonzero(done)[0]:
                ora.reset_one(int(i), cells[i], (xs[i], ys[i]))
        assert np.array_equal(g_obs.cpu().numpy().view(np.uint32), ora.obs.view(np.uint32)), t
    env.check(); env.close()


def test_philox_maps_distribution_matches_reference_generator():
    """Same construction => same distribution: compare the device's maps with maps drawn by the
    Python port of the reference generator (global `random`, plantos_env.py:338-372)."""
    from oracle.plantos_oracle import PlantOSOracle
    from rl_env_b200 import PlantOSVecEnv
    n = 8192
    for kw in (T_KW, DFLT_KW):
        g, p = kw["grid_size"], kw["num_plants"]
        env = PlantOSVecEnv(n, seed=21, **kw)
        env.reset()
        st = env.get_state()
        cells = st["cells"].cpu().numpy()
        xs, ys = st["x"].cpu().numpy(), st["y"].cpu().numpy()
        env.close()
        obst = (cells == 1)
        plants = (cells >= 2)
        assert (plants.sum(axis=(1, 2)) == p).all()
        assert not obst[:, 0, :].any() and not obst[:, -1, :].any() and not obst[:, :, 0].any() and not obst[:, :, -1].any()
        assert (cells[np.arange(n), xs, ys] == 0).all()          # rover on a free, plant-less cell
        assert np.array_equal(st["total_cells"].cpu().numpy(), g * g - obst.sum(axis=(1, 2)))
        assert np.array_equal(st["thirsty_plants"].cpu().numpy(), (cells == 3).sum(axis=(1, 2)))
        random.seed(1)
        ref = PlantOSOracle(**kw)
        m = 1500
        ref_obst = np.zeros((m, g, g), bool)
        ref_thirsty = np.zeros(m)
        ref_rover = np.zeros((m, 2))
        for k in range(m):
            ref.generate_map()
            plane = ref.cell_plane()
            ref_obst[k] = plane == 1
            ref_thirsty[k] = (plane == 3).sum()
            ref_rover[k] = ref.rover_pos
        # obstacle-count distribution: means within 5 standard errors, similar spread
        a, b = obst.sum(axis=(1, 2)).astype(float), ref_obst.sum(axis=(1, 2)).astype(float)
        se = np.sqrt(a.var() / n + b.var() / m)
        assert abs(a.mean() - b.mean()) < 5 * se, (a.mean(), b.mean(), se)
        assert abs(a.std() - b.std()) < 0.15 * b.std() + 0.2
        assert a.min() >= 4 * (kw["num_obstacles"] // 3 > 0) and a.max() <= 9 * (kw["num_obstacles"] // 3)
        # per-cell obstacle marginal
        pa, pb = obst.mean(axis=0), ref_obst.mean(axis=0)
        assert np.abs(pa - pb).max() < 0.06, np.abs(pa - pb).max()
        assert np.corrcoef(pa.ravel(), pb.ravel())[0, 1] > 0.9
        # thirsty fraction 0.7 (plantos_env.py:368) and uniform rover placement
        frac = (cells == 3).sum() / (n * p)
        assert abs(frac - 0.7) < 5 * np.sqrt(0.21 / (n * p)), frac
        assert abs(ref_thirsty.mean() / p - 0.7) < 0.05
        assert abs(xs.mean() - ref_rover[:, 0].mean()) < 5 * np.sqrt(xs.var() / n + ref_rover[:, 0].var() / m)
        assert abs(ys.mean() - ref_rover[:, 1].mean()) < 5 * np.sqrt(ys.var() / n + ref_rover[:, 1].var() / m)


def test_full_size_properties_and_kernel_agreement():
    """131 072 envs (BASELINE.json configs[3] per-GPU share): invariants the domain offers, and
    bit-identical outputs from the two independent kernels on the same seeds and actions."""
    import torch
    from rl_env_b200 import PlantOSVecEnv
    n = 131072
    fast = PlantOSVecEnv(n, kernel="fast", seed=11, max_steps=40, **T_KW)
    gen = PlantOSVecEnv(n, kernel="generic", seed=11, max_steps=40, **T_KW)
    assert fast.kernel_name == "fast" and gen.kernel_name == "generic"
    of, og = fast.reset(), gen.reset()
    assert torch.equal(of, og)
    g = torch.Generator(device="cuda").manual_seed(5)
    valid_moves = torch.zeros(n, dtype=torch.int64, device="cuda")
    for t in range(45):
        actions = torch.randint(0, 5, (n,), device="cuda", generator=g)
        of, rf, df, _ = fast.step(actions)
        og, rg, dg, _ = gen.step(actions)
        assert torch.equal(of, og), f"step {t}: observations differ between kernels"
        assert torch.equal(rf, rg) and torch.equal(df, dg)
        assert torch.equal(fast.terminated, gen.terminated) and torch.equal(fast.truncated, gen.truncated)
        # observation invariants (test_environment.py:183-195 lifted to the batch)
        lid = of[:, :80].view(n, 16, 5)
        assert torch.all(lid[:, :, 1:].sum(dim=2) == 1.0)
        d6 = lid[:, :, 0] * 6
        assert torch.all((d6 - d6.round()).abs() < 1e-5) and d6.min() >= 1 - 1e-5 and d6.max() <= 6 + 1e-5
        assert of.min() >= 0.0 and of.max() <= 1.0
        assert bool(df.any()) == (t == 39), "every env truncates exactly at max_steps"
        valid_moves += ((rf == rf.new_tensor(9.9)) | (rf == rf.new_tensor(-1.1)) | (rf > 50)).long()
        if t == 38:
            st = fast.get_state()
            cells, visits = st["cells"], st["visits"]
            assert torch.all((cells >= 2).sum(dim=(1, 2)) == 10)                        # plants conserved
            assert torch.equal((cells == 3).sum(dim=(1, 2)).int(), st["thirsty_plants"])
            assert torch.equal((visits > 0).sum(dim=(1, 2)).int(), st["explored_cells"])
            assert torch.equal(visits.sum(dim=(1, 2)), 1 + valid_moves)                 # one visit per valid move
            assert torch.equal(st["step_count"], torch.full_like(st["step_count"], 39))
            assert torch.all(cells[torch.arange(n, device="cuda"), st["x"].long(), st["y"].long()] != 1)
    sf, sg = fast.episode_stats(all_reduce=False), gen.episode_stats(all_reduce=False)
    assert sf == sg and sf["episodes"] == n and sf["truncated"] == n and sf["length_sum"] == 40 * n
    fast.check(); gen.check()
    fast.close(); gen.close()


@pytest.mark.parametrize("kernel", ["fast", "generic"])
def test_maze_maps_injected_lockstep(kernel):
    """Maps from the host-side maze generator (oracle.ref_maps, the Gradio fork's 'maze' algorithm:
    corridors and rooms, a much denser LIDAR workload than the cluster maps) pushed as recorded maps:
    the device and the restated reference env agree step by step, auto-resets included."""
    import random
    import torch
    from oracle.plantos_oracle import OracleVecEnv
    from rl_env_b200 import PlantOSVecEnv
    from oracle.ref_maps import make_maps
    n, episodes, steps = 36, 6, 260
    random.seed(12)
    cells, rover = make_maps("maze", n, episodes, T_KW["grid_size"], T_KW["num_plants"], T_KW["num_obstacles"])
    maps = [[(cells[i, e], tuple(int(v) for v in rover[i, e])) for e in range(episodes)] for i in range(n)]
    ora = OracleVecEnv(n, maps=maps, max_steps=60, **T_KW)
    env = PlantOSVecEnv(n, map_source="injected", kernel=kernel, max_steps=60, **T_KW)
    env.push_maps(cells, rover)
    assert env.kernel_name == kernel
    assert np.array_equal(env.reset().cpu().numpy(), ora.reset())
    rng = np.random.default_rng(12)
    for t in range(steps):
        a = rng.integers(0, 5, size=n)
        o_obs, o_rew, o_done, _ = ora.step(a)
        g_obs, g_rew, g_done, _ = env.step(a)
        assert np.array_equal(g_obs.cpu().numpy().view(np.uint32), o_obs.view(np.uint32)), t
        assert np.array_equal(g_rew.cpu().numpy(), o_rew) and np.array_equal(g_done.cpu().numpy(), o_done), t
    env.check()
