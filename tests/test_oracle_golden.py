"""The CPU oracles (Python port and C restatement) must reproduce every trajectory recorded
from the unmodified reference (tests/golden/*.npz, made by tests/golden/make_golden.py)
EXACTLY: bit-equal float32 observations, exact float64 rewards, identical integer state."""
import numpy as np
import pytest

from replay import CURRICULUM_FIXTURES, FIXTURES, COracleBackend, PyOracleBackend, check_replay, load_fixture


@pytest.mark.parametrize("name", FIXTURES)
def test_python_port_replays_reference(name):
    fx = load_fixture(name)
    steps = 1200 if name == "replay_T_8env" else None  # keeps the CPU suite short; C oracle runs it all
    res = check_replay(fx, PyOracleBackend(fx), steps=steps)
    assert res["bitexact_obs"] == 1 and res["bitexact_reward"] == 1
    assert res["steps"] > 0


@pytest.mark.parametrize("name", FIXTURES)
def test_c_oracle_replays_reference(name):
    fx = load_fixture(name)
    res = check_replay(fx, COracleBackend(fx))
    assert res["bitexact_obs"] == 1 and res["bitexact_reward"] == 1
    assert res["episodes"] == len(fx["term_t"])


def test_python_port_rewards_are_exact_doubles():
    """Rewards of the port are the reference's python floats, bit for bit (not just 1e-5)."""
    fx = load_fixture("replay_tiny_4env")
    be = PyOracleBackend(fx)
    be.reset()
    for t in range(300):
        _, rew, _, infos = be.env.step(fx["actions"][t])
        want = fx["rewards"][t]
        got = np.array([be.env._ep_rewards[i][-1] if be.env._ep_rewards[i] else np.nan for i in range(be.n)])
        keep = ~np.isnan(got)
        assert np.array_equal(got[keep], want[keep])


def test_fixture_covers_edge_cases():
    """The fixtures exercise: completion bonus + termination, truncation, hydrated watering
    (the reference's TypeError branch), rover standing on a plant, rays longer than the grid."""
    tiny = load_fixture("replay_tiny_4env")
    assert tiny["terminated"].sum() > 0 and np.isclose(tiny["rewards"], 59.9).any()
    t8 = load_fixture("replay_T_8env")
    assert t8["truncated"].sum() == 24 and int(t8["cfg_mistake_steps"]) > 0
    assert np.isclose(t8["rewards"], -10.1).any()
    assert (t8["lidar_dist"][..., 1] == 1).any()  # ray 1 has offset (0,0) at r=1: sees the rover's own cell
    assert int(tiny["cfg_lidar_range"]) > int(tiny["cfg_grid_size"])


@pytest.mark.parametrize("name", CURRICULUM_FIXTURES)
def test_python_port_replays_reference_curriculum(name):
    """CurriculumOracle (the restated CurriculumWrapper, both variants) against trajectories
    recorded through the reference's own wrapper class: persistent visit counts, threshold
    terminations, threshold increments, the fresh-looking reset observation."""
    fx = load_fixture(name)
    res = check_replay(fx, PyOracleBackend(fx))
    assert res["episodes"] >= 40 and res["bitexact_obs"] == 1 and res["bitexact_reward"] == 1
    assert len(np.unique(fx["cur_threshold"])) >= 5            # thresholds were reached and raised


# ---- observation batches left behind by the reference's own training runs (make_last_obs.py)
def _last_obs():
    import os
    path = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "ref_last_obs.npz")
    return dict(np.load(path))


def test_tables_reproduce_every_value_in_the_reference_runs_last_obs():
    """`_last_obs` of the four saved policies = 256 real observations of the unmodified reference
    (training preset).  Every float in them must be bit-identical to an entry of the host tables the
    kernels read (dist r/R, one-hot 0/1, pos x/G, visits min(k,10)/10): that pins the table
    arithmetic against numbers the reference itself produced."""
    from rl_env_b200 import tables
    G, R, C = 25, 6, 16
    dist = set(tables.distance_table(R).view(np.uint32).tolist())
    pos = set(tables.position_table(G).view(np.uint32).tolist())
    vis = set(tables.visit_table().view(np.uint32).tolist())
    onehot = set(np.array([0.0, 1.0], dtype=np.float32).view(np.uint32).tolist())
    batches = _last_obs()
    assert len(batches) == 4
    seen_dist, seen_vis = set(), set()
    for obs in batches.values():
        assert obs.shape == (64, 5 * C + 27) and obs.dtype == np.float32
        bits = obs.view(np.uint32)
        lidar = bits[:, :5 * C].reshape(64, C, 5)
        assert set(lidar[:, :, 0].ravel().tolist()) <= dist
        assert set(lidar[:, :, 1:].ravel().tolist()) <= onehot
        assert (obs[:, :5 * C].reshape(64, C, 5)[:, :, 1:].sum(axis=2) == 1.0).all()       # valid one-hot
        assert set(bits[:, 5 * C:5 * C + 2].ravel().tolist()) <= pos
        assert set(bits[:, 5 * C + 2:].ravel().tolist()) <= vis
        seen_dist |= set(lidar[:, :, 0].ravel().tolist()); seen_vis |= set(bits[:, 5 * C + 2:].ravel().tolist())
    # the reference runs exercise every distance 1/6 .. 6/6 and every visit level 0 .. 1.0
    assert seen_dist == dist - {np.float32(0.0).view(np.uint32).item()}
    assert seen_vis == vis


def test_oracle_observations_obey_what_the_reference_runs_show():
    """Structural facts of the reference's `_last_obs` that the oracle must share: a ray that reports
    EMPTY is at full range, an out-of-range visit cell reads 1.0, the rover's own cell has been
    visited (centre of the 5x5 window > 0)."""
    from oracle.plantos_oracle import PRESETS, PlantOSOracle
    import random
    C = 16
    def facts(obs):
        lid = obs[:5 * C].reshape(C, 5)
        empty = lid[:, 1] == 1.0
        assert (lid[empty, 0] == 1.0).all()
        assert obs[5 * C + 2 + 12] > 0.0
    for obs_batch in _last_obs().values():
        for row in obs_batch:
            facts(row)
    random.seed(7)
    env = PlantOSOracle(**PRESETS["T"])
    obs = env.reset()
    rng = np.random.default_rng(7)
    for _ in range(300):
        obs, *_ = env.step(int(rng.integers(0, 5)))
        facts(np.asarray(obs, dtype=np.float32))
