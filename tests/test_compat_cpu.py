"""CPU-side checks of the reference call surface (no GPU work): names, signatures, checkpoint and cfg
compatibility.  The numerical behaviour of these classes is tested on the GPU box (test_gpu_compat.py)."""
import configparser
import inspect
import os

import pytest
import torch

from multiagent_gnn_policies_b200 import compat

compat.install()

REF = "/root/reference"


def test_import_surface_and_signatures():
    import gym
    import gym_flock
    from learner.actor import Actor
    from learner.gnn_baseline import train_baseline
    from learner.gnn_cloning import train_cloning
    from learner.gnn_dagger import DAGGER, train_dagger
    from learner.replay_buffer import ReplayBuffer, Transition
    from learner.state_with_delay import MultiAgentStateWithDelay
    assert list(inspect.signature(Actor.__init__).parameters)[1:] == ["n_s", "n_a", "hidden_layers", "k", "ind_agg"]
    assert list(inspect.signature(MultiAgentStateWithDelay.__init__).parameters)[1:] == \
        ["device", "args", "env_state", "prev_state", "k"]
    assert list(inspect.signature(DAGGER.__init__).parameters)[1:] == ["device", "args", "k"]
    for name in ("select_action", "gradient_step", "save_model", "load_model"):
        assert hasattr(DAGGER, name)
    assert Transition._fields == ('state', 'action', 'done', 'next_state', 'reward')
    assert callable(train_dagger) and callable(train_cloning) and callable(train_baseline)
    env = gym.make("FlockingRelative-v0")
    assert isinstance(env.env, gym_flock.envs.FlockingRelativeEnv)
    for name in ("seed", "reset", "step", "render", "close"):
        assert hasattr(env, name)
    for name in ("params_from_cfg", "controller"):
        assert hasattr(env.env, name)
    buf = ReplayBuffer(max_size=3)
    for i in range(5):
        buf.insert(Transition(i, i, i, i, i))
    assert buf.curr_size == 3 and sorted(t.state for t in buf.buffer) == [2, 3, 4]
    assert len(buf.sample(2)) == 2


def test_actor_parameter_names_match_reference_checkpoint_layout():
    from learner.actor import Actor
    a = Actor(6, 2, [32, 32], 3, 0)
    shapes = {k: tuple(v.shape) for k, v in a.state_dict().items()}
    assert shapes == {"conv_layers.0.weight": (32, 6, 3, 1), "conv_layers.0.bias": (32,),
                      "conv_layers.1.weight": (32, 32, 1, 1), "conv_layers.1.bias": (32,),
                      "conv_layers.2.weight": (2, 32, 1, 1), "conv_layers.2.bias": (2,)}
    ckpt = os.path.join(REF, "models", "actor_FlockingRelative-v0_dagger_k3")
    if os.path.exists(ckpt):
        a.load_state_dict(torch.load(ckpt, map_location="cpu"))


def test_actor_refuses_cpu_tensors():
    from learner.actor import Actor
    from multiagent_gnn_policies_b200.engine import FgnnError
    a = Actor(6, 2, [8], 2, 0)
    with pytest.raises(FgnnError):
        a(torch.zeros(1, 2, 6, 5), torch.zeros(1, 2, 5, 5))


@pytest.mark.skipif(not os.path.isdir(REF), reason="reference tree only exists in the build container")
def test_every_flocking_relative_cfg_is_accepted_by_the_env_shim():
    import gym
    import glob
    n = 0
    for path in sorted(glob.glob(os.path.join(REF, "cfg", "*.cfg"))):
        cp = configparser.ConfigParser()
        try:
            cp.read(path)
        except configparser.Error:
            continue            # e.g. cfg/default_baseline.cfg repeats an option: rejected by configparser itself
        for sec in (cp.sections() or [cp.default_section]):
            args = cp[sec]
            if args.get("env") != "FlockingRelative-v0":
                continue
            env = gym.make(args.get("env"))
            env.env.params_from_cfg(args)
            assert env.env.n_agents == args.getint("n_agents")
            assert env.env.comm_radius == args.getfloat("comm_radius")
            n += 1
    assert n > 30


# ---- environment variants: host-side logic only, no engine is created ---------------------

def _args(**kv):
    cp = configparser.ConfigParser()
    cp.read_dict({"DEFAULT": {k: str(v) for k, v in kv.items()}})
    return cp["DEFAULT"]


def test_variant_envs_are_registered_and_configurable():
    import gym
    import gym_flock
    import numpy as np
    names = {"FlockingLeader-v0": gym_flock.envs.FlockingLeaderEnv, "FlockingTwoFlocks-v0": gym_flock.envs.FlockingTwoFlocksEnv,
             "FlockingStochastic-v0": gym_flock.envs.FlockingStochasticEnv}
    for name, cls in names.items():
        env = gym.make(name)
        assert isinstance(env.env, cls) and isinstance(env.env, gym_flock.envs.FlockingRelativeEnv)
    with pytest.raises(KeyError):
        gym.make("FlockingAirsimAccel-v0")              # needs an external simulator: out of scope

    lead = gym.make("FlockingLeader-v0").env
    lead.params_from_cfg(_args(comm_radius=1.0, n_agents=40, v_max=3.0, dt=0.01, k=1))
    assert lead.mask.shape == (40,) and lead.mask[:2].sum() == 0 and lead.mask[2:].all()
    np.random.seed(7)
    x = lead._sample_initial_state()
    assert np.all(x[0:2, 2:4] == x[0, 2])               # the leaders share one scalar velocity

    two = gym.make("FlockingTwoFlocks-v0").env
    two.params_from_cfg(_args(comm_radius=1.0, n_agents=100, v_max=3.0, dt=0.01, k=2))
    np.random.seed(8)
    x = two._sample_initial_state()
    assert x[:50, 0].mean() < 0 < x[50:, 0].mean()

    sto = gym.make("FlockingStochastic-v0").env
    sto.params_from_cfg(_args(comm_radius=1.5, n_agents=100, v_max=0.5, k=4))          # the *_stoch cfgs carry no dt
    assert sto.dt == sto.dt_mean and sto.v_max == 0.5
    np.random.seed(9)
    dts = [sto.draw_dt() for _ in range(200)]
    assert min(dts) >= sto.dt_min and abs(np.mean(dts) - sto.dt_mean) < 4 * sto.dt_sigma / np.sqrt(200)


def test_variant_sampling_matches_the_oracle_draw_for_draw():
    """compat envs draw from the global numpy RNG exactly like the oracle classes draw from a RandomState of the same
    seed: the GPU parity tests of the variants start from identical states."""
    import gym
    import numpy as np
    from oracle import flock_env
    for name, ocls in (("FlockingLeader-v0", flock_env.FlockingLeaderOracle),
                       ("FlockingTwoFlocks-v0", flock_env.FlockingTwoFlocksOracle)):
        env = gym.make(name).env
        env.params_from_cfg(_args(comm_radius=1.0, n_agents=60, v_max=3.0, dt=0.01, k=2))
        np.random.seed(31)
        x = env._sample_initial_state()
        o = ocls(n_agents=60, rng=np.random.RandomState(31))
        o.reset()
        np.testing.assert_array_equal(x, o.x)


@pytest.mark.skipif(not os.path.isdir(REF), reason="reference tree only exists in the build container")
def test_every_variant_cfg_of_the_reference_is_accepted():
    import gym
    import glob
    seen = set()
    for path in sorted(glob.glob(os.path.join(REF, "cfg", "*.cfg"))):
        cp = configparser.ConfigParser()
        try:
            cp.read(path)
        except configparser.Error:
            continue
        for sec in (cp.sections() or [cp.default_section]):
            args = cp[sec]
            name = args.get("env")
            if name in ("FlockingLeader-v0", "FlockingTwoFlocks-v0", "FlockingStochastic-v0"):
                env = gym.make(name)
                env.env.params_from_cfg(args)
                assert env.env.n_agents == args.getint("n_agents") and env.env.dt > 0
                seen.add(name)
    assert seen == {"FlockingLeader-v0", "FlockingTwoFlocks-v0", "FlockingStochastic-v0"}
