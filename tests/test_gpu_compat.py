"""The reference call surface (gym / gym_flock / learner shims) on the GPU: a test_model.py-shaped
rollout, the dense attributes of the state container, the expert controller, and a tiny training run."""
import configparser

import numpy as np
import pytest

from conftest import rel_inf, load_golden
from oracle import flock_env, learner as olearner

pytestmark = pytest.mark.gpu

CFG = """
[DEFAULT]
alg = dagger
batch_size = 4
buffer_size = 200
updates_per_step = 3
seed = 11
actor_lr = 5e-5
n_train_episodes = 2
beta_coeff = 0.993
test_interval = 1
n_test_episodes = 1
k = 3
hidden_size = 32
gamma = 0.99
tau = 0.5
env = FlockingRelative-v0
v_max = 3.0
comm_radius = 1.0
n_agents = 100
n_actions = 2
n_states = 6
debug = True
header = reward
dt = 0.01
"""


def make_args(**over):
    cp = configparser.ConfigParser()
    cp.read_string(CFG)
    for k, v in over.items():
        cp["DEFAULT"][k] = str(v)
    return cp["DEFAULT"]


@pytest.fixture()
def compat():
    from multiagent_gnn_policies_b200 import compat as c
    c.install()
    return c


def test_rollout_like_test_model_py(compat):
    """Same call sequence as the reference's test_model.py:14-47, checked step by step against the oracle."""
    import torch
    import gym
    import gym_flock
    from learner.gnn_dagger import DAGGER
    from learner.state_with_delay import MultiAgentStateWithDelay
    g = load_golden("ckpt_n100_k3")
    args = make_args()
    env = gym.make(args.get('env'))
    assert isinstance(env.env, gym_flock.envs.FlockingRelativeEnv)
    env.env.params_from_cfg(args)
    env.seed(11)
    np.random.seed(11)
    device = torch.device("cuda:0")
    learner = DAGGER(device, args)
    learner.actor.load_state_dict({k: torch.from_numpy(v) for k, v in g["state_dict"].items()})
    layers = olearner.weights_from_state_dict(g["state_dict"])

    state = MultiAgentStateWithDelay(device, args, env.reset(), prev_state=None)
    x = env.env.get_state()
    ostate, total = None, 0.0
    for t in range(6):
        sv, sn, _, _ = flock_env.compute_helpers(x, 1.0)
        ostate = olearner.DelayState((sv, sn), prev_state=ostate, k=3)
        a_ref = olearner.select_action(layers, ostate)
        action = learner.select_action(state)
        assert action.is_cuda and tuple(action.shape) == (100, 2)
        a = action.cpu().numpy()
        assert rel_inf(a, a_ref) <= 1e-5
        # the lazily materialised dense attributes are the reference's tensors
        assert rel_inf(state.delay_state.cpu().numpy(), ostate.delay_state) <= 1e-6
        assert rel_inf(state.delay_gso.cpu().numpy(), ostate.delay_gso) <= 1e-6
        np.testing.assert_array_equal(state.network.cpu().numpy(), ostate.network)
        # dense route through Actor.forward (CUDA dense kernel) agrees too
        with torch.no_grad():
            mu = learner.actor(state.delay_state, state.delay_gso)
        assert rel_inf(mu[0, 0].t().cpu().numpy(), a_ref) <= 1e-5
        next_state, reward, done, _ = env.step(a)
        x = flock_env.integrate(x, a, 0.01)
        np.testing.assert_array_equal(env.env.get_state(), x)
        assert reward == pytest.approx(flock_env.instant_cost(x), rel=1e-9)
        state = MultiAgentStateWithDelay(device, args, next_state, prev_state=state)
        total += reward
    env.render()
    env.close()


def test_time_limit_and_reset_constraints(compat):
    import gym
    args = make_args(n_agents=60)
    env = gym.make('FlockingRelative-v0')
    env.env.params_from_cfg(args)
    np.random.seed(3)
    values, network = env.reset()
    assert values.shape == (60, 6) and network.shape == (60, 60)
    x = env.env.get_state()
    _, r2 = flock_env.pair_terms(x)
    assert np.sqrt(r2.min()) >= 0.1 and (r2 < 1.0).sum(axis=1).min() >= 2
    net = np.asarray(network)
    assert np.sum(np.diag(net)) == 0
    _, sn, _, _ = flock_env.compute_helpers(x, 1.0)
    np.testing.assert_array_equal(net.astype(np.float32), sn.astype(np.float32))
    done, steps = False, 0
    while not done:
        _, _, done, _ = env.step(np.zeros((60, 2), np.float32))
        steps += 1
    assert steps == 200
    env.close()


@pytest.mark.parametrize("R", [1.0, 0.8, 1.5])
@pytest.mark.parametrize("centralized", [True, False])
def test_expert_controller_matches_oracle(compat, R, centralized):
    from multiagent_gnn_policies_b200.engine import FlockEngine
    x = flock_env.synthetic_state(300, seed=5, density=1.6)
    eng = FlockEngine(n_agents=300, comm_radius=R, edge_capacity=64)
    eng.reset(x)
    u = eng.controller(centralized=centralized)
    u_ref = flock_env.controller(x, R, R * R, centralized=centralized)
    np.testing.assert_allclose(u, u_ref, rtol=2e-6, atol=2e-7)
    # batched episodes: every episode gets its own velocity mean and neighbour set
    eng2 = FlockEngine(n_agents=100, n_episodes=3, comm_radius=R, edge_capacity=64)
    eng2.reset(x)
    u2 = eng2.controller(centralized=centralized).reshape(3, 100, 2)
    for b in range(3):
        xb = x[b * 100:(b + 1) * 100]
        np.testing.assert_allclose(u2[b], flock_env.controller(xb, R, R * R, centralized=centralized),
                                   rtol=2e-6, atol=2e-7)
    eng.close()
    eng2.close()


def test_gradient_step_and_tiny_training_run(compat):
    import torch
    import gym
    from learner.gnn_dagger import DAGGER, train_dagger
    from learner.gnn_cloning import train_cloning
    from learner.gnn_baseline import train_baseline
    from learner.replay_buffer import Transition
    from learner.state_with_delay import MultiAgentStateWithDelay
    args = make_args(n_agents=40, actor_lr=1e-3)
    device = torch.device("cuda:0")
    np.random.seed(0)
    torch.manual_seed(0)
    env = gym.make('FlockingRelative-v0')
    env.env.params_from_cfg(args)
    env._max_episode_steps = 12
    learner = DAGGER(device, args)
    # a fixed batch: the loss must go down under repeated gradient steps
    state = MultiAgentStateWithDelay(device, args, env.reset(), prev_state=None)
    trans = []
    for _ in range(6):
        u = env.env.controller()
        nxt, r, d, _ = env.step(u)
        nstate = MultiAgentStateWithDelay(device, args, nxt, prev_state=state)
        label = torch.Tensor(u).to(device).transpose(1, 0).reshape((1, 1, 2, 40))
        trans.append(Transition(state, label, None, nstate, r))
        state = nstate
    batch = Transition(*zip(*trans))
    losses = [learner.gradient_step(batch) for _ in range(30)]
    assert losses[-1] < losses[0]
    # after the update the engine-backed select_action uses the NEW weights
    a_sparse = learner.select_action(state).cpu().numpy()
    with torch.no_grad():
        a_dense = learner.actor(state.delay_state, state.delay_gso)[0, 0].t().cpu().numpy()
    assert rel_inf(a_sparse, a_dense) <= 1e-5
    env.close()
    for train in (train_dagger, train_cloning):
        env = gym.make('FlockingRelative-v0')
        env.env.params_from_cfg(args)
        env._max_episode_steps = 8
        stats = train(env, args, device)
        assert np.isfinite(stats['mean'])
    env = gym.make('FlockingRelative-v0')
    args_b = make_args(n_agents=40, centralized=True, n_test_episodes=2)
    env.env.params_from_cfg(args_b)
    env._max_episode_steps = 8
    stats = train_baseline(env, args_b)
    assert np.isfinite(stats['mean'])
