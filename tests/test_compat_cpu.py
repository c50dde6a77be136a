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
