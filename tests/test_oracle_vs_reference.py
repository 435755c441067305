"""Pin the oracle port against the UNMODIFIED reference, run side by side (build container
only: the reference checkout is not shipped to the GPU box, where this module is skipped and
the committed golden vectors take over)."""
import random

import numpy as np
import pytest

from oracle.ref_shim import reference_available

pytestmark = pytest.mark.skipif(not reference_available(), reason="reference checkout not mounted")


@pytest.mark.parametrize("preset,seed,steps", [("T", 0, 1500), ("T", 3, 1200), ("DFLT", 1, 1500), ("XL", 0, 150)])
def test_port_equals_reference_step_by_step(preset, seed, steps):
    from oracle.plantos_oracle import PRESETS, PlantOSOracle
    from oracle.ref_shim import ReferenceEnv
    kw = PRESETS[preset]
    ref, ora = ReferenceEnv(**kw), PlantOSOracle(**kw)
    random.seed(seed)
    o1, i1 = ref.reset()
    random.seed(seed)
    o2, i2 = ora.reset()
    assert np.array_equal(o1.view(np.uint32), o2.view(np.uint32)) and o1.dtype == o2.dtype == np.float32
    assert ref.env.obstacles == ora.obstacles and list(ref.env.plants.items()) == list(ora.plants.items())
    assert ref.env.rover_pos == ora.rover_pos and set(i1) == set(i2)
    rng = np.random.default_rng(seed)
    for t in range(steps):
        a = int(rng.integers(5))
        r1, r2 = ref.step(a), ora.step(a)
        assert np.array_equal(r1[0].view(np.uint32), r2[0].view(np.uint32)), t
        assert r1[1] == r2[1] and r1[2] == r2[2] and r1[3] == r2[3], t
        assert r1[4] == r2[4], t
        assert np.array_equal(ref.env.visit_counts, ora.visit_counts)
        assert np.array_equal(ref.env.explored_map, ora.explored_map)
        if r1[2] or r1[3]:
            state = random.getstate()
            ref.reset()
            random.setstate(state)
            ora.reset()
            assert ref.env.obstacles == ora.obstacles and list(ref.env.plants.items()) == list(ora.plants.items())
            assert ref.env.rover_pos == ora.rover_pos
    assert ref.mistake_steps == ora.mistake_steps


def test_reference_bug_is_what_the_policy_says():
    """plantos_env.py:213-222: watering a hydrated plant raises TypeError after step_count was
    incremented and with nothing else changed; the shim substitutes the documented -10."""
    from oracle.ref_shim import ReferenceEnv, load_reference
    mod = load_reference()
    env = mod.PlantOSEnv(grid_size=7, num_plants=3, num_obstacles=0)
    random.seed(0)
    env.reset()
    pos = next(iter(env.plants))
    env.plants[pos] = False
    env.rover_pos = pos
    before = (env.step_count, dict(env.plants), env.visit_counts.copy())
    with pytest.raises(TypeError):
        env.step(4)
    assert env.step_count == before[0] + 1 and env.plants == before[1]
    assert np.array_equal(env.visit_counts, before[2])
    shim = ReferenceEnv(grid_size=7, num_plants=3, num_obstacles=0)
    random.seed(0)
    shim.reset()
    shim.env.plants[pos] = False
    shim.env.rover_pos = pos
    _, reward, _, _, _ = shim.step(4)
    assert reward == -0.1 + -10 and shim.mistake_steps == 1


def test_golden_fixtures_are_reproducible():
    """Regenerating a fixture from the reference gives the committed bytes' content."""
    import importlib.util, os, tempfile
    here = os.path.dirname(os.path.abspath(__file__))
    spec = importlib.util.spec_from_file_location("make_golden", os.path.join(here, "golden", "make_golden.py"))
    mg = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mg)
    with tempfile.TemporaryDirectory() as tmp:
        mg.OUT = tmp
        name, n, steps, seed, kw, ms = mg.FIXTURES[3]
        mg.record(name, n, steps, seed, kw, max_steps=ms)
        new = np.load(os.path.join(tmp, name + ".npz"))
        old = np.load(os.path.join(here, "golden", name + ".npz"))
        assert set(new.files) == set(old.files)
        for k in old.files:
            assert np.array_equal(new[k], old[k]), k


@pytest.mark.parametrize("variant,init,max_eps", [("a2c", 40.0, 3), ("dqn", 30.0, 4)])
def test_curriculum_port_matches_the_reference_wrapper_class(variant, init, max_eps):
    """CurriculumOracle vs the reference's own CurriculumWrapper (class compiled from
    A2C_training.py:37-109 / trainingCode.py:24-98) around the real env, in lock-step: observations,
    rewards, flags, visit_counts, thresholds -- including the observation returned by reset."""
    import random
    from oracle.plantos_oracle import CurriculumOracle, PlantOSOracle
    from oracle.ref_shim import ReferenceEnv, load_curriculum_wrapper

    def cells_of(env):
        g = env.grid_size
        p = np.zeros((g, g), dtype=np.uint8)
        for (x, y) in env.obstacles:
            p[x, y] = 1
        for (x, y), thirsty in env.plants.items():
            p[x, y] = 3 if thirsty else 2
        return p

    kw = dict(grid_size=7, num_plants=3, num_obstacles=4, lidar_range=4, lidar_channels=8)
    random.seed(3)
    np.random.seed(3)
    ref = load_curriculum_wrapper(variant)(ReferenceEnv(**kw), initial_threshold=init, max_threshold=100.0)
    ref.max_episodes_per_maze = max_eps
    ora = CurriculumOracle(PlantOSOracle(**kw), variant, max_episodes_per_maze=max_eps)
    raw = ref.env.env
    raw.max_steps = ora.env.max_steps = 90
    o, _ = ref.reset()
    o2, _ = ora.reset(cells_of(raw), raw.rover_pos)
    assert np.array_equal(o, o2)
    rng = np.random.default_rng(5)
    episodes, thresholds = 0, set()
    for t in range(4000):
        a = int(rng.integers(0, 5))
        o, r, te, tr, _ = ref.step(a)
        o2, r2, te2, tr2, _ = ora.step(a)
        assert np.array_equal(o, o2) and r == r2 and te == te2 and tr == tr2, t
        assert np.array_equal(raw.visit_counts, ora.env.visit_counts)
        assert ref.exploration_threshold == ora.exploration_threshold and ref.maze_completed == ora.maze_completed
        thresholds.add(ref.exploration_threshold)
        if te or tr:
            episodes += 1
            o, _ = ref.reset()
            o2, _ = ora.reset(cells_of(raw), raw.rover_pos)
            assert np.array_equal(o, o2) and np.array_equal(raw.visit_counts, ora.env.visit_counts)
    assert episodes >= 40 and len(thresholds) >= 4


@pytest.mark.parametrize("algo", ["maze", "original"])
def test_host_map_generators_equal_the_gradio_fork(algo):
    """oracle.ref_maps (host-side generators for map injection) against `_generate_map` of the Gradio
    fork's env class (gradio-app/plantos_env_new.py:353-604) with the same `random.seed`: same obstacle /
    plant / rover cells and the same number of RNG draws."""
    import importlib, os, sys, types
    from oracle import ref_shim
    from oracle.ref_maps import maze_map, original_map
    fork_dir = os.path.join(ref_shim.REFERENCE_DIR, "gradio-app")
    if not os.path.isfile(os.path.join(fork_dir, "plantos_env_new.py")):
        pytest.skip("fork not in the checkout")
    ref_shim._install_stubs()
    sys.modules["pygame"].Surface = type("Surface", (), {})
    for name, attrs in (("plantos_utils", ["print_reset_info", "print_step_info", "print_episode_summary"]),
                        ("plantos_3d_viewer_new", ["PlantOS3DViewer"])):
        if name not in sys.modules:
            m = types.ModuleType(name)
            for a in attrs:
                setattr(m, a, (lambda *args, **kw: None) if a.startswith("print") else type(a, (), {}))
            sys.modules[name] = m
    if fork_dir not in sys.path:
        sys.path.insert(0, fork_dir)
    fork = importlib.import_module("plantos_env_new")
    gen = {"maze": maze_map, "original": original_map}[algo]
    for g, pl, ob in ((25, 10, 12), (21, 8, 50), (31, 12, 30), (13, 4, 9), (8, 3, 6)):
        for seed in range(5):
            env = fork.PlantOSEnvNew(grid_size=g, num_plants=pl, num_obstacles=ob, lidar_range=4, lidar_channels=8,
                                     observation_mode="lidar", map_generation_algo=algo)
            random.seed(seed)
            env.obstacles, env.plants = set(), {}
            env._generate_map()
            after_ref = random.random()
            random.seed(seed)
            cells, rover = gen(g, pl, ob)
            after_ours = random.random()
            want = np.zeros((g, g), np.uint8)
            for (x, y) in env.obstacles:
                want[x, y] = 1
            for (x, y), thirsty in env.plants.items():
                want[x, y] = 3 if thirsty else 2
            assert np.array_equal(cells, want) and tuple(rover) == tuple(env.rover_pos) and after_ref == after_ours
