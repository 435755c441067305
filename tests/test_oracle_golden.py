"""The CPU oracles (Python port and C restatement) must reproduce every trajectory recorded
from the unmodified reference (tests/golden/*.npz, made by tests/golden/make_golden.py)
EXACTLY: bit-equal float32 observations, exact float64 rewards, identical integer state."""
import numpy as np
import pytest

from replay import FIXTURES, COracleBackend, PyOracleBackend, check_replay, load_fixture


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
