"""rl_env_b200.sb3: the SB3 `VecEnv` subclass over the batched simulator.  stable_baselines3 is not part of the
image, so the class is built over a stand-in base with SB3's constructor signature
(VecEnv.__init__(num_envs, observation_space, action_space)); the CPU test drives the adaptor over a scripted
inner env, the GPU test over the real simulator against the restated DummyVecEnv + Monitor."""
import numpy as np
import pytest
import torch


class StubVecEnvBase:
    """stable_baselines3.common.vec_env.VecEnv, as far as a subclass relies on it."""

    def __init__(self, num_envs, observation_space, action_space):
        self.num_envs, self.observation_space, self.action_space = num_envs, observation_space, action_space

    def step(self, actions):                       # SB3: step = step_async + step_wait
        self.step_async(actions)
        return self.step_wait()


class ScriptedEnv:
    """The PlantOSVecEnv surface the adaptor touches, on CPU tensors."""

    def __init__(self, n=3, d=5):
        self.num_envs, self.obs_dim, self.device = n, d, torch.device("cpu")
        self.observation_space, self.action_space = ("box", d), ("discrete", 5)
        self.t, self.closed, self.attrs = 0, False, {}

    def reset(self):
        self.t = 0
        return torch.zeros((self.num_envs, self.obs_dim))

    def step_async(self, actions):
        assert actions.dtype == torch.int64 and actions.shape == (self.num_envs,)
        self.last = actions.clone()

    def step_wait(self):
        self.t += 1
        obs = torch.full((self.num_envs, self.obs_dim), float(self.t)) + self.last[:, None].float()
        done = torch.tensor([self.t % 2 == 0, False, self.t % 3 == 0])
        infos = [{"step_count": self.t, "TimeLimit.truncated": bool(done[i]),
                  **({"terminal_observation": torch.full((self.obs_dim,), -1.0), "episode": {"r": 1.5, "l": self.t, "t": 0.0}}
                     if done[i] else {})} for i in range(self.num_envs)]
        return obs, torch.arange(self.num_envs).float() * self.t, done, infos

    def close(self):
        self.closed = True

    def get_attr(self, name, indices=None):
        return [self.attrs.get(name, 7)] * self.num_envs

    def set_attr(self, name, value, indices=None):
        self.attrs[name] = value

    def env_method(self, name, *a, indices=None, **k):
        return [name] * self.num_envs

    def env_is_wrapped(self, cls, indices=None):
        return [False] * self.num_envs

    def seed(self, seed=None):
        return [None] * self.num_envs


def test_adaptor_over_a_scripted_env():
    from rl_env_b200.sb3 import vecenv_class
    inner = ScriptedEnv()
    venv = vecenv_class(StubVecEnvBase)(inner)
    assert venv.num_envs == 3 and venv.observation_space == ("box", 5)
    obs = venv.reset()
    assert isinstance(obs, np.ndarray) and obs.shape == (3, 5) and obs.dtype == np.float32
    for t in range(1, 7):
        obs, rew, done, infos = venv.step(np.array([0, 1, 4]))
        assert obs.dtype == np.float32 and rew.dtype == np.float32 and done.dtype == np.bool_
        assert np.array_equal(obs[:, 0], t + np.array([0, 1, 4], np.float32))
        assert np.array_equal(rew, np.arange(3, dtype=np.float32) * t)
        assert list(done) == [t % 2 == 0, False, t % 3 == 0]
        assert isinstance(infos, list) and len(infos) == 3
        for i in range(3):
            if done[i]:                            # "done" mode: dicts only where SB3 reads them
                assert isinstance(infos[i]["terminal_observation"], np.ndarray)
                assert infos[i]["episode"]["l"] == t and infos[i]["TimeLimit.truncated"]
            else:
                assert infos[i] == {}
    with pytest.raises(RuntimeError):
        venv.step_wait()
    venv.set_attr("max_steps", 50)
    assert venv.get_attr("max_steps") == [50, 50, 50]
    assert venv.env_method("close") == ["close"] * 3 and venv.env_is_wrapped(object) == [False] * 3
    full = vecenv_class(StubVecEnvBase)(ScriptedEnv(), info_mode="full")
    full.reset()
    _, _, _, infos = full.step(np.zeros(3, np.int64))
    assert all(info["step_count"] == 1 for info in infos)
    venv.close()
    assert inner.closed


def test_make_sb3_vecenv_needs_sb3():
    from rl_env_b200.sb3 import make_sb3_vecenv
    try:
        import stable_baselines3  # noqa: F401
    except Exception:
        with pytest.raises(ImportError):
            make_sb3_vecenv(4)


@pytest.mark.gpu
def test_adaptor_matches_dummyvecenv_monitor_on_the_device():
    """The adaptor over the real simulator, stepped with numpy actions like an SB3 algorithm does, against the
    restated DummyVecEnv + Monitor of the reference on the tiny fixture (terminations and truncations)."""
    from replay import PyOracleBackend, fixture_kwargs, load_fixture
    from rl_env_b200 import PlantOSVecEnv
    from rl_env_b200.sb3 import vecenv_class
    fx = load_fixture("replay_tiny_4env")
    ora = PyOracleBackend(fx).env
    env = PlantOSVecEnv(4, map_source="injected", max_steps=int(fx["cfg_max_steps"]), full_infos=False,
                        info_keywords=("exploration_percentage",), **fixture_kwargs(fx))
    env.push_maps(fx["maps_cells"], fx["maps_rover"])
    venv = vecenv_class(StubVecEnvBase)(env)
    assert np.array_equal(venv.reset(), ora.reset())
    ndone = 0
    for t in range(600):
        o_obs, o_rew, o_done, o_infos = ora.step(fx["actions"][t])
        g_obs, g_rew, g_done, g_infos = venv.step(fx["actions"][t])
        assert np.array_equal(g_obs, o_obs) and np.array_equal(g_rew, o_rew) and np.array_equal(g_done, o_done)
        for i in range(4):
            if o_done[i]:
                ndone += 1
                gi, oi = g_infos[i], o_infos[i]
                assert gi["episode"]["r"] == oi["episode"]["r"] and gi["episode"]["l"] == oi["episode"]["l"]
                assert gi["TimeLimit.truncated"] == oi["TimeLimit.truncated"]
                assert np.array_equal(gi["terminal_observation"], oi["terminal_observation"])
                assert gi["episode"]["exploration_percentage"] == oi["exploration_percentage"]   # Monitor(info_keywords=...)
            else:
                assert g_infos[i] == {}
    assert ndone >= 2
    # max_steps is a plain attribute of the reference env (plantos_env.py:120): changing it takes effect
    venv.set_attr("max_steps", 3)
    seen = np.zeros(4, bool)
    for t in range(600, 612):
        _, _, done, _ = venv.step(fx["actions"][t])
        seen |= done
    assert seen.all()                                          # every env truncates within the new 3-step horizon
    venv.close()
