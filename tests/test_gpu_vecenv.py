"""The SB3-facing surface: infos, Monitor episode records, episode statistics, spaces."""
import numpy as np
import pytest

from replay import PyOracleBackend, fixture_kwargs, load_fixture

pytestmark = pytest.mark.gpu


def test_infos_match_dummyvecenv_monitor():
    """Step the restated DummyVecEnv+Monitor (oracle) and PlantOSVecEnv side by side on the
    tiny fixture (terminations + truncations) and compare the info dicts key by key."""
    from rl_env_b200 import PlantOSVecEnv
    fx = load_fixture("replay_tiny_4env")
    ora = PyOracleBackend(fx).env
    env = PlantOSVecEnv(4, map_source="injected", max_steps=int(fx["cfg_max_steps"]), **fixture_kwargs(fx))
    env.push_maps(fx["maps_cells"], fx["maps_rover"])
    assert env.full_infos
    assert env.observation_space.shape == (fx["obs"].shape[2],) and env.action_space.n == 5
    o_obs, g_obs = ora.reset(), env.reset().cpu().numpy()
    assert np.array_equal(o_obs, g_obs)
    seen_done = 0
    for t in range(700):
        o_obs, o_rew, o_done, o_infos = ora.step(fx["actions"][t])
        g_obs, g_rew, g_done, g_infos = env.step(fx["actions"][t])
        assert np.array_equal(g_obs.cpu().numpy(), o_obs)
        assert np.array_equal(g_rew.cpu().numpy(), o_rew)
        assert np.array_equal(g_done.cpu().numpy(), o_done)
        assert isinstance(g_infos, list) and len(g_infos) == 4
        for i in range(4):
            gi, oi = g_infos[i], o_infos[i]
            for key in ("rover_position", "thirsty_plants", "hydrated_plants", "total_plants", "step_count",
                        "explored_cells", "total_cells", "exploration_percentage", "lidar_range",
                        "lidar_channels", "collided_with_wall", "total_collisions", "TimeLimit.truncated"):
                assert gi[key] == oi[key], (t, i, key, gi[key], oi[key])
            assert ("episode" in gi) == ("episode" in oi) == bool(o_done[i])
            if o_done[i]:
                seen_done += 1
                assert gi["episode"]["r"] == oi["episode"]["r"] and gi["episode"]["l"] == oi["episode"]["l"]
                assert np.array_equal(gi["terminal_observation"].cpu().numpy(), oi["terminal_observation"])
    assert seen_done >= 2
    stats = env.episode_stats(all_reduce=False)
    k = fx["term_t"] < 700
    assert stats["episodes"] == k.sum()
    assert abs(stats["return_sum"] - fx["ep_r"][k].sum()) < 1e-6
    assert stats["length_sum"] == fx["ep_l"][k].sum()
    assert stats["terminated"] == fx["terminated"][:700].sum()
    assert stats["truncated"] == fx["truncated"][:700].sum()
    env.close()


def test_lazy_infos_and_state_roundtrip():
    from rl_env_b200 import LazyInfos, PlantOSVecEnv
    env = PlantOSVecEnv(256, seed=2, grid_size=25, num_plants=10, num_obstacles=12, lidar_range=6, lidar_channels=16)
    env.reset()
    rng = np.random.default_rng(0)
    for _ in range(20):
        obs, rew, done, infos = env.step(rng.integers(0, 5, 256))
    assert isinstance(infos, LazyInfos) and len(infos) == 256
    info = infos[17]
    assert info["step_count"] == 20 and info["total_plants"] == 10
    # get_state -> set_state into a second simulator reproduces the trajectory (MCTS-style copy,
    # mcts_custom_trainer.py:218-243)
    st = env.get_state()
    twin = PlantOSVecEnv(256, seed=99, grid_size=25, num_plants=10, num_obstacles=12, lidar_range=6, lidar_channels=16)
    twin.reset()
    twin.set_state(st["cells"], st["visits"], {k: st[k] for k in
                   ("x", "y", "step_count", "explored_cells", "total_collisions", "collided_with_wall",
                    "completion_bonus_given", "episode", "watered")})
    for _ in range(30):
        a = rng.integers(0, 5, 256)
        o1, r1, d1, _ = env.step(a)
        o2, r2, d2, _ = twin.step(a)
        assert np.array_equal(o1.cpu().numpy(), o2.cpu().numpy())
        assert np.array_equal(r1.cpu().numpy(), r2.cpu().numpy())
    env.close(); twin.close()


def test_step_host_matches_device_step():
    from rl_env_b200 import PlantOSVecEnv
    kw = dict(grid_size=21, num_plants=8, num_obstacles=50, lidar_range=2, lidar_channels=10)
    a = PlantOSVecEnv(513, seed=5, **kw)
    b = PlantOSVecEnv(513, seed=5, **kw)
    a.reset(); b.reset()
    rng = np.random.default_rng(1)
    for _ in range(25):
        act = rng.integers(0, 5, 513)
        o1, r1, d1, _ = a.step(act)
        o2, r2, d2 = b.step_host(act)
        assert np.array_equal(o1.cpu().numpy(), o2) and np.array_equal(r1.cpu().numpy(), r2)
        assert np.array_equal(d1.cpu().numpy(), d2)
    a.close(); b.close()


@pytest.mark.parametrize("kernel", ["fast", "generic"])
def test_episode_log_and_monitor_csv_match_the_oracle_monitor(kernel, tmp_path):
    """Device episode log (one entry per finished episode, appended inside the step kernel) against
    the restated DummyVecEnv + Monitor: same (env, r, l) multiset, r equal after Monitor's own
    round(sum, 6); and the monitor.csv files carry the reference's format
    (train_improved1/gym/env_0.monitor.csv)."""
    import json
    from rl_env_b200 import MonitorCSV, PlantOSVecEnv
    name = "replay_T_8env" if kernel == "fast" else "replay_tiny_4env"
    fx = load_fixture(name)
    n = fx["actions"].shape[1]
    ora = PyOracleBackend(fx).env
    env = PlantOSVecEnv(n, map_source="injected", kernel=kernel, max_steps=int(fx["cfg_max_steps"]), **fixture_kwargs(fx))
    env.push_maps(fx["maps_cells"], fx["maps_rover"])
    mon = MonitorCSV(env, str(tmp_path), per_env_files=True)
    ora.reset(); env.reset()
    want = []
    steps = min(2200, fx["actions"].shape[0])
    for t in range(steps):
        _, _, o_done, o_infos = ora.step(fx["actions"][t])
        env.step(fx["actions"][t])
        for i in range(n):
            if o_done[i]:
                want.append((i, o_infos[i]["episode"]["l"], o_infos[i]["episode"]["r"], t))
        if t % 700 == 699:
            mon.flush()
    mon.close()
    assert mon.dropped == 0 and len(want) >= 2
    rows = {}
    for i in range(n):
        path = tmp_path / f"env_{i}.monitor.csv"
        if not path.exists():
            continue
        lines = path.read_text().splitlines()
        head = json.loads(lines[0][1:])
        assert lines[0].startswith("#") and set(head) == {"t_start", "env_id"} and lines[1] == "r,l,t"
        rows[i] = [(float(a), int(b), float(c)) for a, b, c in (ln.split(",") for ln in lines[2:])]
    got = sorted((i, l, r) for i, rs in rows.items() for r, l, _ in rs)
    assert got == sorted((i, l, round(r, 6)) for i, l, r, _ in want)
    # raw log entries: exact doubles, step numbers, flags
    env2 = PlantOSVecEnv(n, map_source="injected", kernel=kernel, max_steps=int(fx["cfg_max_steps"]), **fixture_kwargs(fx))
    env2.push_maps(fx["maps_cells"], fx["maps_rover"])
    env2.enable_episode_log(capacity=3)                      # tiny capacity: overflow is counted, not fatal
    env2.reset()
    for t in range(steps):
        env2.step(fx["actions"][t])
    eps, dropped = env2.drain_episode_log()
    assert len(eps) == min(3, len(want)) and dropped == len(want) - len(eps)
    for e in eps:
        match = [w for w in want if w[0] == int(e["env"]) and w[3] == int(e["step_seq"])]
        assert match and match[0][1] == int(e["l"]) and match[0][2] == round(float(e["r"]), 6)
        assert int(e["flags"]) in (1, 2, 3)


@pytest.mark.parametrize("pipelined", [True, False])
@pytest.mark.parametrize("n,kernel", [(4096, "fast"), (8192, "fast"), (515, "fast"), (300, "generic")])
def test_graph_rollout_equals_single_steps(n, kernel, pipelined):
    """make_rollout(K): K steps captured in one CUDA graph (incl. the programmatic-dependent-launch
    edges of the fast kernel) give exactly what K single steps give, replay after replay."""
    import torch
    from rl_env_b200 import PlantOSVecEnv, PRESETS
    kw = dict(PRESETS["training"], max_steps=37, seed=11, kernel=kernel, full_infos=False)
    a, b = PlantOSVecEnv(n, **kw), PlantOSVecEnv(n, **kw)
    assert torch.equal(a.reset(), b.reset())
    K = 12
    roll = b.make_rollout(K, pipelined=pipelined)   # pipelined: steps ordered tile by tile on the device (no-op with a ragged tail / generic)
    g = torch.Generator(device="cuda"); g.manual_seed(3)
    for rep in range(5):                                   # 60 steps: every env auto-resets at least once
        acts = torch.randint(0, 5, (K, n), device="cuda", generator=g)
        obs_k, rew_k, done_k = roll(acts)
        for t in range(K):
            obs, rew, done, _ = a.step(acts[t])
            assert torch.equal(obs, obs_k[t]) and torch.equal(rew, rew_k[t]) and torch.equal(done, done_k[t]), (rep, t)
    sa, sb = a.get_state(), b.get_state()
    assert all(torch.equal(sa[k], sb[k]) for k in sa)
    a.check(); b.check()
    a.close(); b.close()


@pytest.mark.parametrize("n,kernel,preset,cur", [(4096, "fast", "training", None), (1027, "fast", "training", None),
                                                 (2048, "fast", "default", None), (96, "fast", "training", "a2c"),
                                                 (200, "generic", "training", None)])
def test_step_many_equals_single_steps(n, kernel, preset, cur):
    """plantos_rollout / env.step_many: K steps in one call (one launch of the state-resident multi-step
    kernel on the fast presets) give exactly what K single steps give -- observations, rewards, flags,
    terminal observations, episode statistics and the full state afterwards -- with auto-resets inside the
    rollouts, a ragged tail (1027), the curriculum wrapper, and single steps mixed in between."""
    import torch
    from rl_env_b200 import PlantOSVecEnv, PRESETS
    kw = dict(PRESETS[preset], max_steps=23, seed=21, kernel=kernel, full_infos=False, curriculum=cur)
    a, b = PlantOSVecEnv(n, **kw), PlantOSVecEnv(n, **kw)
    assert torch.equal(a.reset(), b.reset())
    g = torch.Generator(device="cuda"); g.manual_seed(8)
    for rep, K in enumerate((7, 1, 16, 5, 16, 16)):
        acts = torch.randint(0, 5, (K, n), device="cuda", generator=g)
        obs_k, rew_k, done_k, term_k, trunc_k = b.step_many(acts, with_flags=True)
        for t in range(K):
            obs, rew, done, _ = a.step(acts[t])
            assert torch.equal(obs, obs_k[t]), (rep, t)
            assert torch.equal(rew, rew_k[t]) and torch.equal(done, done_k[t]), (rep, t)
            assert torch.equal(a.terminated, term_k[t]) and torch.equal(a.truncated, trunc_k[t]), (rep, t)
        assert torch.equal(a.terminal_observation, b.terminal_observation)
        if rep == 2:                                       # single steps between rollouts
            x = torch.randint(0, 5, (n,), device="cuda", generator=g)
            oa, ra, da, _ = a.step(x)
            ob, rb, db, _ = b.step(x)
            assert torch.equal(oa, ob) and torch.equal(ra, rb) and torch.equal(da, db)
    if kernel == "fast":
        assert b.last_step_kernel == "k_rollout_tile"
    sa, sb = a.get_state(), b.get_state()
    assert all(torch.equal(sa[k], sb[k]) for k in sa)
    assert a.episode_stats() == b.episode_stats()
    a.check(); b.check()
    a.close(); b.close()


def test_monitor_info_keywords_is_watering_and_set_attr(tmp_path):
    """SB3 Monitor(info_keywords=...) columns from the device episode log (the reference's EvaluationCallback
    wants exploration_percentage per episode, A2C_training.py:161-179), the Gradio fork's info['is_watering']
    (gradio-app/plantos_env_new.py:184), and set_attr for the attributes the reference lets callers change
    (max_steps, plantos_env.py:120; the reward constants, :76-83)."""
    import torch
    from rl_env_b200 import MonitorCSV, PlantOSVecEnv
    fx = load_fixture("replay_tiny_4env")
    ora = PyOracleBackend(fx).env
    env = PlantOSVecEnv(4, map_source="injected", max_steps=int(fx["cfg_max_steps"]), **fixture_kwargs(fx))
    env.push_maps(fx["maps_cells"], fx["maps_rover"])
    mon = MonitorCSV(env, str(tmp_path), info_keywords=("exploration_percentage", "total_collisions"))
    with pytest.raises(ValueError):
        MonitorCSV(env, str(tmp_path / "x"), info_keywords=("rover_position",))
    ora.reset(); env.reset()
    want = []
    for t in range(700):
        _, _, o_done, o_infos = ora.step(fx["actions"][t])
        _, _, _, g_infos = env.step(fx["actions"][t])
        for i in range(4):
            assert g_infos[i]["is_watering"] == bool(fx["actions"][t][i] >= 4)
            if o_done[i]:
                want.append((o_infos[i]["episode"]["l"], o_infos[i]["exploration_percentage"], o_infos[i]["total_collisions"]))
    mon.close()
    lines = (tmp_path / "monitor.csv").read_text().splitlines()
    assert lines[1] == "r,l,t,exploration_percentage,total_collisions"
    got = sorted((int(l), float(x), int(c)) for _, l, _, x, c in (ln.split(",") for ln in lines[2:]))
    assert got == sorted((l, float(x), c) for l, x, c in want) and len(got) >= 2
    # set_attr: a new reward constant and a new horizon apply to the following steps
    env.set_attr("R_INVALID", -7.0)
    env.set_attr("max_steps", 2)
    assert env.get_attr("max_steps") == [2] * 4
    _, rew, d1, _ = env.step(np.zeros(4, np.int64))
    d1 = d1.clone()
    _, rew, d2, _ = env.step(np.zeros(4, np.int64))
    assert bool((d1 | d2).all())                           # every env hits the new horizon within two steps
    with pytest.raises(AttributeError):
        env.set_attr("grid_size", 30)
    twin = PlantOSVecEnv(64, seed=1, rewards={"r_invalid": -7.0}, **fixture_kwargs(fx))
    late = PlantOSVecEnv(64, seed=1, **fixture_kwargs(fx))
    late.set_attr("R_INVALID", -7.0)
    twin.reset(); late.reset()
    g = torch.Generator(device="cuda"); g.manual_seed(0)
    for _ in range(30):
        a = torch.randint(0, 5, (64,), device="cuda", generator=g)
        assert torch.equal(twin.step(a)[1], late.step(a)[1])
    env.close(); twin.close(); late.close()


def test_curriculum_reuse_map_keeps_the_maze():
    """reuse_map=True: while the wrapper stays on a maze, every reset regenerates exactly that maze (what
    A2C_training.py:75-86 `reset(seed=self.current_maze_seed)` means to do); the default draws a new map at
    every reset, like the reference really does.  With 3 episodes per maze and a threshold no random walk
    reaches in 12 steps, episodes 0-1 share a maze, 2-4 share the next, 5-7 the one after."""
    import torch
    from rl_env_b200 import PlantOSVecEnv, PRESETS
    from rl_env_b200.vec_env import CURRICULA
    n, horizon = 96, 12
    for reuse in (True, False):
        env = PlantOSVecEnv(n, seed=9, max_steps=horizon, kernel="fast", full_infos=False,
                            curriculum=dict(CURRICULA["a2c"], initial_threshold=99.0, reuse_map=reuse), **PRESETS["training"])
        env.reset()
        maps = []
        g = torch.Generator(device="cuda"); g.manual_seed(1)
        for ep in range(8):
            st = env.get_state()
            maps.append((st["cells"].clone(), st["x"].clone(), st["y"].clone()))
            for _ in range(horizon):
                _, _, done, _ = env.step(torch.randint(0, 4, (n,), device="cuda", generator=g))   # moves only: plants stay as generated
            assert bool(done.all())
        same = lambda a, b: all(torch.equal(u, v) for u, v in zip(maps[a], maps[b]))
        if reuse:
            assert same(0, 1) and same(2, 3) and same(2, 4) and same(5, 6) and same(5, 7)
            assert not same(0, 2) and not same(2, 5)
        else:
            assert not any(same(a, a + 1) for a in range(7))
        env.check(); env.close()


def test_pipelined_eager_steps_equal_plain_steps():
    """plantos_set_pipelining on eager launches: back-to-back step_async calls into an observation ring
    overlap on the device (per-tile counters order them) and give exactly the plain results; every
    observation of the ring is checked after the burst."""
    import torch
    from rl_env_b200 import PlantOSVecEnv, PRESETS
    n, ring, steps = 16384, 6, 90
    kw = dict(PRESETS["training"], max_steps=29, seed=5, kernel="fast", full_infos=False)
    a, b = PlantOSVecEnv(n, **kw), PlantOSVecEnv(n, obs_ring=ring, **kw)
    b.set_pipelining(True)
    assert torch.equal(a.reset(), b.reset())
    g = torch.Generator(device="cuda"); g.manual_seed(4)
    acts = torch.randint(0, 5, (steps, n), device="cuda", generator=g)
    want = []
    for t in range(steps):
        obs, rew, done, _ = a.step(acts[t])
        want.append((obs.clone(), rew.clone(), done.clone()))
    for t0 in range(0, steps, ring - 1):                   # bursts of ring-1 steps without a host sync
        got = []
        for t in range(t0, min(steps, t0 + ring - 1)):
            b.step_async(acts[t])
            obs, rew, done, _ = b.step_wait()
            got.append((t, obs))
        torch.cuda.synchronize()
        for t, obs in got:
            assert torch.equal(obs, want[t][0]), t
        assert torch.equal(b._rewards, want[t][1]) and torch.equal(b._dones, want[t][2])
    assert b.last_step_kernel == "k_step_tile"
    sa, sb = a.get_state(), b.get_state()
    assert all(torch.equal(sa[k], sb[k]) for k in sa)
    a.check(); b.check()
    a.close(); b.close()


@pytest.mark.parametrize("n,preset,cur", [(16384, "training", None), (8192, "default", None), (4096, "training", "a2c")])
def test_pipelined_rollouts_equal_plain_steps(n, preset, cur):
    """plantos_set_pipelining on step_many: consecutive multi-step launches (same output buffers, mixed with
    single steps into an observation ring) overlap on the device through the per-tile counters and give exactly
    what plain single steps give, auto-resets (and the curriculum wrapper) included."""
    import torch
    from rl_env_b200 import PlantOSVecEnv, PRESETS
    ring = 4
    kw = dict(PRESETS[preset], max_steps=23, seed=9, kernel="fast", full_infos=False, curriculum=cur)
    a, b = PlantOSVecEnv(n, **kw), PlantOSVecEnv(n, obs_ring=ring, **kw)
    b.set_pipelining(True)
    assert torch.equal(a.reset(), b.reset())
    g = torch.Generator(device="cuda"); g.manual_seed(11)
    plan = [5, 5, 1, 1, 7, 5, 1, 12, 12, 3]                 # K of every call (1 = step_async / step_wait)
    acts = torch.randint(0, 5, (sum(plan), n), device="cuda", generator=g)
    want = []
    for t in range(acts.shape[0]):
        obs, rew, done, _ = a.step(acts[t])
        want.append((obs.clone(), rew.clone(), done.clone()))
    t, checks = 0, []
    for k in plan:                                          # everything enqueued back to back, no host sync
        if k == 1:
            b.step_async(acts[t])
            obs, rew, done, _ = b.step_wait()
            checks.append((t, obs, None, None, 0))
        else:
            obs, rew, done = b.step_many(acts[t:t + k])
            # (the next call with the same K reuses these buffers: copy them out, stream-ordered)
            checks.append((t, obs.clone(), rew.clone(), done.clone(), k))
        t += k
    torch.cuda.synchronize()
    for t0, obs, rew, done, k in checks:
        if k == 0:
            assert torch.equal(obs, want[t0][0]), t0
        else:
            for j in range(k):
                assert torch.equal(obs[j], want[t0 + j][0]), (t0, j)
                assert torch.equal(rew[j], want[t0 + j][1]) and torch.equal(done[j], want[t0 + j][2]), (t0, j)
    assert b.last_step_kernel == "k_rollout_tile"
    sa, sb = a.get_state(), b.get_state()
    assert all(torch.equal(sa[k], sb[k]) for k in sa)
    assert a.episode_stats(all_reduce=False) == b.episode_stats(all_reduce=False)
    a.check(); b.check()
    a.close(); b.close()


def test_pipelined_rollouts_full_size_equal_plain_rollouts():
    """BASELINE's per-GPU share (131 072 envs, staggered episode phases as in bench.py): 96 steps as six
    pipelined 16-step launches into the SAME buffers against the same launches one at a time -- last
    observations, rewards, flags, full state and episode statistics identical."""
    import torch
    from rl_env_b200 import PlantOSVecEnv, PRESETS
    n = 131072
    kw = dict(PRESETS["training"], seed=3, kernel="fast", full_infos=False)
    a, b = PlantOSVecEnv(n, **kw), PlantOSVecEnv(n, **kw)
    b.set_pipelining(True)
    assert torch.equal(a.reset(), b.reset())
    gid = torch.arange(n, device="cuda", dtype=torch.int64)
    phase = (((gid * 2654435761) % 4294967296) % 1000).to(torch.int32)
    a.set_state(scalars={"step_count": phase}); b.set_state(scalars={"step_count": phase})
    g = torch.Generator(device="cuda"); g.manual_seed(2)
    acts = torch.randint(0, 5, (96, n), device="cuda", generator=g)
    for c in range(6):                                      # b: six launches back to back, no host sync in between
        outs_b = b.step_many(acts[16 * c:16 * c + 16], with_flags=True)
    for c in range(6):
        outs_a = a.step_many(acts[16 * c:16 * c + 16], with_flags=True)
        torch.cuda.synchronize()
    torch.cuda.synchronize()
    for x, y in zip(outs_a, outs_b):
        assert torch.equal(x, y)
    assert torch.equal(a.terminal_observation, b.terminal_observation)
    sa, sb = a.get_state(), b.get_state()
    assert all(torch.equal(sa[k], sb[k]) for k in sa)
    sta, stb = a.episode_stats(all_reduce=False), b.episode_stats(all_reduce=False)
    assert sta == stb and sta["episodes"] > 10000
    a.check(); b.check()
    a.close(); b.close()


def test_rollout_policy_matches_the_mcts_heuristic():
    """plantos_rollout_policy against the restated MCTS rollout policy (mcts_custom_trainer.py:168-216)
    on the same states and the same injected uniforms, while both sides follow that policy for 300
    steps (visit counts beyond the 4-bit range included: tiny grid, long episodes)."""
    import torch
    from oracle.plantos_oracle import rollout_policy
    from rl_env_b200 import PlantOSVecEnv
    from oracle.plantos_oracle import OracleVecEnv
    fx = load_fixture("replay_tiny_4env")
    n = fx["actions"].shape[1]
    # the exploring policy finishes the tiny grid quickly: cycle the fixture's maps 40 times
    k = int(fx["n_maps"].min())
    cells, rover = np.tile(fx["maps_cells"][:, :k], (1, 40, 1, 1)), np.tile(fx["maps_rover"][:, :k], (1, 40, 1))
    maps = [[(cells[i, e], tuple(rover[i, e])) for e in range(cells.shape[1])] for i in range(n)]
    ora = OracleVecEnv(n, maps=maps, max_steps=int(fx["cfg_max_steps"]), **fixture_kwargs(fx))
    env = PlantOSVecEnv(n, map_source="injected", max_steps=int(fx["cfg_max_steps"]), **fixture_kwargs(fx))
    env.push_maps(cells, rover)
    ora.reset(); env.reset()
    g = torch.Generator(device="cuda"); g.manual_seed(5)
    kinds = set()
    for t in range(600):
        u = torch.rand((n, 2), dtype=torch.float32, device="cuda", generator=g)
        acts = env.rollout_policy(u)
        uh = u.cpu().numpy()
        want = [rollout_policy(ora.envs[i], uh[i, 0], uh[i, 1]) for i in range(n)]
        assert acts.cpu().tolist() == want, t
        kinds |= {("h" if uh[i, 0] < 0.7 else "r") for i in range(n)}
        o_obs, _, _, _ = ora.step(want)
        g_obs, _, _, _ = env.step(acts)
        assert np.array_equal(g_obs.cpu().numpy(), o_obs)
    assert kinds == {"h", "r"}
    # visit counts beyond the 4-bit range (overflow plane): overwrite them on both sides
    gsz = int(fx["cfg_grid_size"])
    big = np.random.default_rng(3).integers(0, 60, size=(n, gsz, gsz)).astype(np.int32)
    env.set_state(visits=torch.as_tensor(big))
    for i in range(n):
        ora.envs[i].visit_counts[:, :] = big[i]
    for t in range(20):
        u = torch.rand((n, 2), dtype=torch.float32, device="cuda", generator=g)
        uh = u.cpu().numpy()
        assert env.rollout_policy(u).cpu().tolist() == [rollout_policy(ora.envs[i], uh[i, 0], uh[i, 1]) for i in range(n)]


def test_more_than_255_plants_are_rejected():
    """The env record counts thirsty plants in 8 bits: recorded maps and injected state beyond that are refused
    (plantos_push_maps: EINVAL; plantos_set_state: flagged on the device, raised by check())."""
    import torch
    from rl_env_b200 import PlantOSVecEnv
    from rl_env_b200._native import PlantOSError
    g = 21
    env = PlantOSVecEnv(2, map_source="injected", grid_size=g, num_plants=8, num_obstacles=0, lidar_range=2, lidar_channels=10)
    cells = np.zeros((2, 1, g, g), dtype=np.uint8)
    cells[:, 0, 1:15, :] = 3                                   # 294 thirsty plants
    rover = np.zeros((2, 1, 2), dtype=np.int16)
    with pytest.raises(PlantOSError, match="255"):
        env.push_maps(cells, rover)
    cells[:, 0, 13:15, :] = 0                                  # 252: accepted
    env.push_maps(cells, rover)
    env.reset()
    env.check()
    too_many = torch.as_tensor(np.ascontiguousarray(cells[:, 0]), device=env.device).clone()
    too_many[:, 13:16, :] = 3
    env.set_state(cells=too_many)
    with pytest.raises(PlantOSError, match="255"):
        env.check()
    env.close()
