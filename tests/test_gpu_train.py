"""Native gradient step (fgnn_trainer_step) against the vectors frozen from the reference's
own DAGGER.gradient_step (learner/gnn_dagger.py:76-96; tests/golden/train_*.npz), against torch autograd at a
size the golden files do not reach, and through the compat DAGGER on engine-recorded states."""
import configparser
import copy

import numpy as np
import pytest

from conftest import rel_inf

pytestmark = pytest.mark.gpu

TOL_LOSS = 1e-5          # relative
TOL_GRAD = 2e-5          # relative to the largest entry of the tensor (fp32 sums in a different order)


def _tensors(sd, device):
    import torch
    n = len(sd) // 2
    out = []
    for i in range(n):
        out.append(torch.tensor(sd[f"conv_layers.{i}.weight"], device=device).contiguous())
        out.append(torch.tensor(sd[f"conv_layers.{i}.bias"], device=device).contiguous())
    return out


def _z_rows(z_bkfn, device):
    """golden z is (B,K,F,N) (actor.py:70); the trainer takes (B,K,N,6) -- the engine's aggregated layout."""
    import torch
    return torch.tensor(np.ascontiguousarray(z_bkfn.transpose(0, 1, 3, 2)), device=device)


def test_loss_and_gradients_match_reference(train_golden):
    import torch
    from multiagent_gnn_policies_b200.engine import ActorTrainer
    g = train_golden
    dev = torch.device("cuda:0")
    tr = ActorTrainer(g["k"], g["hidden"], g["n_layers"])
    params = _tensors(g["sd0"], dev)
    before = [p.clone() for p in params]
    loss, grads = tr.step(_z_rows(g["z"][0], dev), torch.tensor(g["target"][0], device=dev), params, apply=False,
                          want_grads=True)
    assert abs(loss.item() - g["loss"][0]) <= TOL_LOSS * abs(g["loss"][0])
    grads = grads.cpu().numpy()
    off = 0
    for i in range(g["n_layers"] + 1):
        for kind in ("weight", "bias"):
            ref = g["grad1"][f"conv_layers.{i}.{kind}"]
            got = grads[off:off + ref.size].reshape(ref.shape)
            off += ref.size
            assert rel_inf(got, ref) <= TOL_GRAD, (i, kind)
    assert off == tr.n_params
    for p, b in zip(params, before):            # apply=False leaves the parameters alone
        assert torch.equal(p, b)
    assert tr.launch_count() == 2
    tr.close()


def test_adam_trajectory_matches_reference(train_golden):
    import torch
    from multiagent_gnn_policies_b200.engine import ActorTrainer
    g = train_golden
    dev = torch.device("cuda:0")
    tr = ActorTrainer(g["k"], g["hidden"], g["n_layers"])
    params = _tensors(g["sd0"], dev)
    m = [torch.zeros_like(p) for p in params]
    v = [torch.zeros_like(p) for p in params]
    lr = g["lr"]
    losses = []
    for s in range(g["steps"]):
        loss, _ = tr.step(_z_rows(g["z"][s], dev), torch.tensor(g["target"][s], device=dev), params, m, v, step=s + 1, lr=lr)
        losses.append(loss.item())
        if s == 0:
            # one Adam step moves every parameter by ~lr: agree with the reference to a small fraction of that
            for i in range(g["n_layers"] + 1):
                ref_w = g["sd1"][f"conv_layers.{i}.weight"]
                assert np.abs(params[2 * i].cpu().numpy() - ref_w).max() <= 2e-3 * lr + 1e-7 * np.abs(ref_w).max()
                assert np.abs(params[2 * i + 1].cpu().numpy() - g["sd1"][f"conv_layers.{i}.bias"]).max() <= 2e-3 * lr + 1e-7
    np.testing.assert_allclose(losses, g["loss"], rtol=5e-5)
    for i in range(g["n_layers"] + 1):
        ref_w = g["sdT"][f"conv_layers.{i}.weight"]
        tol = 1e-2 * lr * g["steps"]
        assert np.abs(params[2 * i].cpu().numpy() - ref_w).max() <= tol + 1e-7 * np.abs(ref_w).max()
        assert np.abs(params[2 * i + 1].cpu().numpy() - g["sdT"][f"conv_layers.{i}.bias"]).max() <= tol + 1e-7
    tr.close()


@pytest.mark.parametrize("batch,n,k,hidden,layers", [(64, 1000, 3, 32, 2), (3, 77, 2, 20, 1), (9, 513, 4, 128, 3)])
def test_gradients_match_torch_autograd_at_scale(batch, n, k, hidden, layers):
    """Sizes the golden files do not reach (many tiles per CTA, ragged last tile, odd widths): a torch fp32
    autograd evaluation of the same readout on the GPU is the reference here."""
    import torch
    from multiagent_gnn_policies_b200.engine import ActorTrainer
    torch.backends.cudnn.allow_tf32 = False            # the torch side must really be fp32 (conv defaults to TF32)
    torch.backends.cuda.matmul.allow_tf32 = False
    dev = torch.device("cuda:0")
    gen = torch.Generator(device="cpu").manual_seed(5)
    dims = [6 * k] + [hidden] * layers + [2]
    params = []
    for i in range(layers + 1):
        shape = (dims[i + 1], 6, k, 1) if i == 0 else (dims[i + 1], dims[i], 1, 1)
        params.append((torch.randn(shape, generator=gen) / np.sqrt(dims[i])).to(dev).requires_grad_(True))
        params.append((0.1 * torch.randn(dims[i + 1], generator=gen)).to(dev).requires_grad_(True))
    z = torch.randn((batch, k, n, 6), generator=gen).to(dev)
    y = torch.randn((batch, 1, 2, n), generator=gen).to(dev)
    x = z.permute(0, 3, 1, 2)                                    # (B,K,N,6) -> (B,F,K,N): actor.py:65
    for i in range(layers + 1):
        x = torch.nn.functional.conv2d(x, params[2 * i], params[2 * i + 1], stride=(k if i == 0 else 1, 1))
        if i < layers:
            x = torch.tanh(x)
    loss_ref = torch.nn.functional.mse_loss(x.view(batch, 1, 2, n), y)
    loss_ref.backward()
    tr = ActorTrainer(k, hidden, layers)
    loss, grads = tr.step(z, y, [p.data for p in params], apply=False, want_grads=True)
    assert abs(loss.item() - loss_ref.item()) <= 2e-5 * abs(loss_ref.item())
    off = 0
    for p in params:
        ref = p.grad.flatten()
        got = grads[off:off + ref.numel()]
        off += ref.numel()
        assert (got - ref).abs().max().item() <= 1e-4 * ref.abs().max().item()
    tr.close()


CFG = """
[DEFAULT]
alg = dagger
batch_size = 6
buffer_size = 200
updates_per_step = 3
seed = 11
actor_lr = 1e-3
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
debug = False
dt = 0.01
"""


def test_compat_gradient_step_native_equals_autograd_path():
    """DAGGER.gradient_step on engine-recorded states (native kernels, sparse replay) against the same update
    through the dense delay_gso / torch-autograd path of an identical learner."""
    import torch
    from multiagent_gnn_policies_b200 import compat
    compat.install()
    import gym
    from learner.gnn_dagger import DAGGER
    from learner.replay_buffer import Transition
    from learner.state_with_delay import MultiAgentStateWithDelay
    cp = configparser.ConfigParser()
    cp.read_string(CFG)
    args = cp["DEFAULT"]
    device = torch.device("cuda:0")
    np.random.seed(3)
    torch.manual_seed(3)
    env = gym.make("FlockingRelative-v0")
    env.env.params_from_cfg(args)
    env.env.record_aggregated = True
    native = DAGGER(device, args)
    dense = DAGGER(device, args)
    dense.actor.load_state_dict(copy.deepcopy(native.actor.state_dict()))
    state = MultiAgentStateWithDelay(device, args, env.reset(), prev_state=None)
    states, labels = [], []
    for _ in range(8):
        u = env.env.controller(False)
        labels.append(torch.Tensor(u).to(device).transpose(1, 0).reshape((1, 1, 2, 100)))
        states.append(state)
        nxt, _, _, _ = env.step(u)
        state = MultiAgentStateWithDelay(device, args, nxt, prev_state=state)
    assert all(s.aggregated is not None and tuple(s.aggregated.shape) == (3, 100, 6) for s in states)
    # the aggregated features are the reference's z = delay_state @ delay_gso (actor.py:70)
    for s in states:
        z_ref = torch.matmul(s.delay_state, s.delay_gso)[0].permute(0, 2, 1)       # (K,N,F)
        assert (s.aggregated - z_ref).abs().max().item() <= 2e-6 * max(z_ref.abs().max().item(), 1.0)

    class Plain:                    # same tensors, no aggregated attribute -> autograd path
        def __init__(self, s):
            self.delay_gso, self.delay_state = s.delay_gso, s.delay_state

    for it in range(3):
        idx = [(it * 3 + j) % 8 for j in range(6)]
        b_native = Transition(tuple(states[i] for i in idx), tuple(labels[i] for i in idx), None, None, None)
        b_dense = Transition(tuple(Plain(states[i]) for i in idx), tuple(labels[i] for i in idx), None, None, None)
        assert native._native_supported(b_native) and not dense._native_supported(b_dense)
        ln = native.gradient_step(b_native)
        ld = dense.gradient_step(b_dense)
        assert abs(ln - ld) <= 2e-5 * abs(ld)
    for pn, pd in zip(native.actor.parameters(), dense.actor.parameters()):
        assert (pn - pd).abs().max().item() <= 3e-2 * 1e-3
    assert native._trainer.launch_count() == 6
    # the torch optimizer state was advanced in place
    st = native.actor_optim.state[next(iter(native.actor.parameters()))]
    assert int(st['step'].item()) == 3 and st['exp_avg'].abs().max().item() > 0
    # the rollout engine sees the updated weights
    a_engine = native.select_action(state)
    a_dense = dense.select_action(Plain(state))
    assert (a_engine - a_dense).abs().max().item() <= 1e-3 * max(a_dense.abs().max().item(), 1.0)
    env.close()


def test_train_dagger_runs_natively():
    """train_dagger end to end on the shims (learner/gnn_dagger.py:126-243): the updates go through the native trainer."""
    import torch
    from multiagent_gnn_policies_b200 import compat
    compat.install()
    import gym
    from learner import gnn_dagger
    cp = configparser.ConfigParser()
    cp.read_string(CFG)
    args = cp["DEFAULT"]
    np.random.seed(1)
    torch.manual_seed(1)
    env = gym.make("FlockingRelative-v0")
    env._max_episode_steps = 12
    env.env.params_from_cfg(args)
    made = []
    orig = gnn_dagger.DAGGER

    class Spy(orig):
        def __init__(self, *a, **k):
            super().__init__(*a, **k)
            made.append(self)

    gnn_dagger.DAGGER = Spy
    try:
        stats = gnn_dagger.train_dagger(env, args, torch.device("cuda:0"))
    finally:
        gnn_dagger.DAGGER = orig
    assert np.isfinite(stats['mean'])
    assert made and made[0]._trainer is not None and made[0]._trainer.launch_count() == 2 * 3 * 2


def test_device_dagger_parallel_episodes():
    """DAGGER for B block-diagonal episodes on the device (BASELINE config C3 in miniature): stored states are the
    engine's aggregated features, sampled batches index (step, episode) correctly, and imitation reduces the loss."""
    import torch
    from multiagent_gnn_policies_b200.dagger import DeviceDagger
    from oracle import flock_env, train as otrain, learner as olearner
    B, N = 4, 60
    rng = np.random.RandomState(2)
    x0 = np.concatenate([flock_env.FlockingRelativeOracle(n_agents=N, rng=rng).sample_initial_state() for _ in range(B)])
    dg = DeviceDagger(N, B, k=3, hidden=32, n_layers=2, lr=2e-3, buffer_steps=32, batch_size=8, seed=5)
    w_before = [p.clone() for p in dg.params]
    ret, loss_sum = dg.run_episode(x0, steps=12, updates=0)
    assert dg.filled == 12 and np.isfinite(ret)
    # the stored label of step 0 is the oracle's expert action for every episode
    for e in range(B):
        u_ref = flock_env.controller(x0[e * N:(e + 1) * N], 1.0, 1.0, centralized=False)
        got = dg.label_buf[0, e * N:(e + 1) * N].cpu().numpy()
        assert np.abs(got - u_ref).max() <= 1e-6
    # a sampled batch is made of whole (step, episode) states
    gen_state = dg.gen.get_state()
    z, y = dg.sample_batch()
    dg.gen.set_state(gen_state)
    pick = torch.randint(0, dg.filled * B, (dg.batch_size,), generator=dg.gen, device=dg.device).cpu().numpy()
    for i, pk in enumerate(pick):
        st, ep = pk // B, pk % B
        np.testing.assert_array_equal(z[i].cpu().numpy(), dg.z_buf[st, :, ep * N:(ep + 1) * N].cpu().numpy())
        np.testing.assert_array_equal(y[i].cpu().numpy(), dg.label_buf[st, ep * N:(ep + 1) * N].cpu().numpy().T)
    # the loss of the native step on that batch equals the oracle's on the same numbers
    layers = olearner.weights_from_state_dict({k_: v.cpu().numpy() for k_, v in dg.state_dict().items()})
    loss_ref, _ = otrain.loss_and_grads(layers, z.cpu().numpy().transpose(0, 1, 3, 2), y.cpu().numpy().reshape(-1, 1, 2, N))
    loss, _ = dg.trainer.step(z, y, dg.params, apply=False)
    assert abs(loss.item() - loss_ref) <= 2e-5 * abs(loss_ref)
    losses = [dg.gradient_step() for _ in range(300)]
    assert np.mean(losses[-20:]) < 0.7 * np.mean(losses[:20])
    assert any((a - b).abs().max().item() > 0 for a, b in zip(dg.params, w_before))
    dg.close()
