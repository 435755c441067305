"""Property tests of the two CPU oracles against each other (hypothesis): random shapes -- rays longer
than the grid, odd ray counts, tiny grids that get fully explored, short episodes -- with the Python
port (pinned against the reference) and the C restatement stepped in lock-step on the same maps and
actions.  The C oracle is what the large GPU lock-step tests compare with, so it must agree with the
port everywhere, not only on the golden shapes."""
import random

import numpy as np
from hypothesis import HealthCheck, given, settings
from hypothesis import strategies as st

from oracle.c_oracle import COracle
from oracle.plantos_oracle import OracleVecEnv


@st.composite
def shapes(draw):
    g = draw(st.integers(5, 14))
    r = draw(st.integers(1, 9))
    c = draw(st.integers(1, 16))
    o = draw(st.integers(0, min(12, (g - 4) * 3)))
    p = draw(st.integers(1, 4))
    max_steps = draw(st.integers(5, 60))
    seed = draw(st.integers(0, 10_000))
    return g, p, o, r, c, max_steps, seed


@settings(max_examples=25, deadline=None, suppress_health_check=[HealthCheck.too_slow])
@given(shapes())
def test_c_oracle_equals_python_port_on_random_shapes(shape):
    g, p, o, r, c, max_steps, seed = shape
    n, steps = 3, 150
    kw = dict(grid_size=g, num_plants=p, num_obstacles=o, lidar_range=r, lidar_channels=c)
    random.seed(seed)
    py = OracleVecEnv(n, max_steps=max_steps, **kw)           # maps from the reference generator's port
    co = COracle(n, g, p, o, r, c, max_steps=max_steps)
    obs = py.reset()
    for i in range(n):
        cells, rover = py.map_log[i][-1]
        co.reset_one(i, cells, rover)
    assert np.array_equal(co.obs.view(np.uint32), obs.view(np.uint32))
    rng = np.random.default_rng(seed)
    consumed = [1] * n
    for t in range(steps):
        a = rng.integers(0, 5, size=n)
        p_obs, p_rew, p_done, p_infos = py.step(a)
        c_obs, c_rew, c_term, c_trunc = co.step(a)
        for i in range(n):
            assert bool(c_term[i]) == p_infos[i]["terminated"] and bool(c_trunc[i]) == p_infos[i]["truncated"], (t, i)
            assert np.float32(c_rew[i]) == p_rew[i], (t, i, c_rew[i], p_rew[i])
            if p_done[i]:
                assert np.array_equal(co.terminal_obs[i].view(np.uint32),
                                      p_infos[i]["terminal_observation"].view(np.uint32)), (t, i)
                assert round(float(co.ep_return[i]), 6) == p_infos[i]["episode"]["r"]
                assert int(co.ep_len[i]) == p_infos[i]["episode"]["l"]
                cells, rover = py.map_log[i][consumed[i]]     # the map the port's auto-reset drew
                consumed[i] += 1
                co.reset_one(i, cells, rover)
        assert np.array_equal(co.obs.view(np.uint32), p_obs.view(np.uint32)), t
    cells, visits, sc = co.get_state()
    for i in range(n):
        st_py = py.envs[i].export_state()
        assert np.array_equal(cells[i], st_py["cells"]) and np.array_equal(visits[i], st_py["visits"])
