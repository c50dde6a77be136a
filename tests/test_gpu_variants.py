"""FlockingLeader / FlockingTwoFlocks / FlockingStochastic on the CUDA engine against
oracle.flock_env's restatement of the same semantics: bit-exact integrated state and degrees, features to fp32
rounding; the leader mask also through the fused closed-loop kernel and the CUDA-graph rollout."""
import configparser

import numpy as np
import pytest

from conftest import load_golden
from oracle import flock_env, learner as olearner

pytestmark = pytest.mark.gpu


def make_args(env_name, **over):
    d = dict(alg="dagger", k="3", hidden_size="32", v_max="3.0", comm_radius="1.0", n_agents="100", n_actions="2",
             n_states="6", dt="0.01", env=env_name)
    d.update({k: str(v) for k, v in over.items()})
    cp = configparser.ConfigParser()
    cp.read_dict({"DEFAULT": d})
    return cp["DEFAULT"]


@pytest.fixture()
def gym_mod():
    from multiagent_gnn_policies_b200 import compat
    compat.install()
    import gym
    import gym_flock  # noqa: F401
    return gym


def _check_obs(obs, x, R):
    sv, sn, _, deg = flock_env.compute_helpers(x, R * R)
    got = np.asarray(obs[0])
    assert np.abs(got - sv.astype(np.float32)).max() <= 2e-7 * max(np.abs(sv).max(), 1.0)
    np.testing.assert_array_equal((np.asarray(obs[1]) != 0).sum(axis=1), deg)


def test_leader_env_matches_oracle(gym_mod):
    args = make_args("FlockingLeader-v0")
    env = gym_mod.make("FlockingLeader-v0")
    env.env.params_from_cfg(args)
    np.random.seed(21)
    obs = env.reset()
    x = env.env.get_state()
    assert np.all(x[0:2, 2:4] == x[0, 2])
    oracle = flock_env.FlockingLeaderOracle(n_agents=100)
    oracle.reset(x)
    _check_obs(obs, x, 1.0)
    rng = np.random.RandomState(0)
    for t in range(6):
        u = rng.uniform(-1, 1, size=(100, 2)).astype(np.float32)
        obs, r, done, _ = env.step(u)
        (_, _), r_ref, _, _ = oracle.step(u)
        np.testing.assert_array_equal(env.env.get_state(), oracle.x)
        assert r == pytest.approx(r_ref, rel=1e-12)
        _check_obs(obs, oracle.x, 1.0)
    np.testing.assert_array_equal(env.env.get_state()[0:2, 2:4], x[0:2, 2:4])
    env.close()


def test_leader_mask_in_the_fused_closed_loop_and_graph_rollout():
    from multiagent_gnn_policies_b200.engine import FlockEngine
    g = load_golden("ckpt_n100_k3")
    layers = olearner.weights_from_state_dict(g["state_dict"])
    x0 = g["x"][0].copy()
    mask = np.ones(100)
    mask[[0, 1, 17]] = 0
    eng = FlockEngine(n_agents=100, k=3, hidden=32, n_layers=2)
    eng.load_state_dict(g["state_dict"])
    eng.set_agent_mask(mask)
    eng.reset(x0)
    x, state = x0, None
    for t in range(4):                                   # fused step kernel, checked against the oracle every step
        sv, sn, _, _ = flock_env.compute_helpers(x, 1.0)
        state = olearner.DelayState((sv, sn), prev_state=state, k=3)
        a = np.empty((100, 2), np.float32)
        eng.step(a, None)
        x = flock_env.integrate(x, a, 0.01, mask=mask)
        np.testing.assert_array_equal(eng.get_state(), x)
    eng.reset(x0)
    eng.rollout(25)                                      # CUDA-graph replay
    xg = eng.get_state()
    np.testing.assert_array_equal(xg[[0, 1, 17], 2:4], x0[[0, 1, 17], 2:4])
    np.testing.assert_array_equal(xg[[0, 1, 17], 0:2], _drift(x0[[0, 1, 17]], 25, 0.01))
    eng.set_agent_mask(None)                             # removing the mask re-captures the graph
    eng.reset(x0)
    eng.rollout(25)
    assert not np.array_equal(eng.get_state()[[0, 1, 17], 2:4], x0[[0, 1, 17], 2:4])
    eng.close()


def _drift(x, steps, dt):
    p = x[:, 0:2].copy()
    for _ in range(steps):
        p = p + x[:, 2:4] * dt
    return p


def test_two_flocks_env_steps_like_the_oracle(gym_mod):
    args = make_args("FlockingTwoFlocks-v0", k=2)
    env = gym_mod.make("FlockingTwoFlocks-v0")
    env.env.params_from_cfg(args)
    np.random.seed(8)
    obs = env.reset()
    x = env.env.get_state()
    assert x[:50, 0].mean() < 0 < x[50:, 0].mean()
    _check_obs(obs, x, 1.0)
    oracle = flock_env.FlockingTwoFlocksOracle(n_agents=100)
    oracle.reset(x)
    for t in range(5):
        u = env.env.controller(False)
        u_ref = oracle.controller(False)
        assert np.abs(u - u_ref).max() <= 1e-6
        assert u.dtype == np.float64                  # the expert's action stays float64 through env.step (gnn_dagger.py:156-163)
        obs, r, _, _ = env.step(u)
        oracle.step(u)
        np.testing.assert_array_equal(env.env.get_state(), oracle.x)
        _check_obs(obs, oracle.x, 1.0)
    env.close()


def test_stochastic_env_uses_a_fresh_dt_every_step(gym_mod):
    args = make_args("FlockingStochastic-v0", v_max=0.5, comm_radius=1.5)
    del args["dt"]                                        # the *_stoch cfgs carry no dt
    env = gym_mod.make("FlockingStochastic-v0")
    env.env.params_from_cfg(args)
    np.random.seed(5)
    obs = env.reset()
    x = env.env.get_state()
    _check_obs(obs, x, 1.5)
    dts = []
    for t in range(6):
        u = env.env.controller(False).astype(np.float32)
        obs, r, _, _ = env.step(u)
        dts.append(env.env.dt)
        x = flock_env.integrate(x, u, env.env.dt)
        np.testing.assert_array_equal(env.env.get_state(), x)
        assert r == pytest.approx(flock_env.instant_cost(x), rel=1e-12)
        _check_obs(obs, x, 1.5)
    assert len(set(dts)) == 6
    env.close()
