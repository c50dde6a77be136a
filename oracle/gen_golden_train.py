"""Generate tests/golden/train_*.npz by running the UNMODIFIED reference ``DAGGER.gradient_step``.

TEST INFRASTRUCTURE.  Run in the build container only (needs /root/reference):

    python -m oracle.gen_golden_train

The reference package ``learner`` is imported from /root/reference (learner/gnn_dagger.py, which pulls
learner/actor.py, learner/state_with_delay.py, learner/replay_buffer.py).  States come from a short rollout
of oracle.flock_env driven by the expert controller (what train_dagger does while beta ~ 1,
gnn_dagger.py:156-163); labels are the expert actions in the reference's (1,1,nA,N) layout
(gnn_dagger.py:173-175).  Frozen per case: initial parameters, the batches (aggregated features z =
delay_state @ delay_gso, actor.py:70, and labels), the loss of every step, the gradients of the first
step, and the parameters after the first and the last step.
"""
import configparser
import os
import sys

import numpy as np
import torch

from oracle import flock_env

REF = os.environ.get("FGNN_REFERENCE", "/root/reference")
OUT = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden")
CKPT = "models/actor_FlockingRelative-v0_dagger_k3"


def make_args(n_agents, k, hidden, n_layers, lr):
    cp = configparser.ConfigParser()
    cp.read_dict({"DEFAULT": {"n_states": "6", "n_actions": "2", "n_agents": str(n_agents), "k": str(k),
                              "hidden_size": str(hidden), "n_layers": str(n_layers), "gamma": "0.99", "tau": "0.1",
                              "actor_lr": repr(lr)}})
    return cp["DEFAULT"]


def run_case(name, n_agents, k, hidden, n_layers, batch, steps, lr, seed, checkpoint=None, comm_radius=1.0,
             v_max=3.0, pool=40):
    sys.path.insert(0, REF)
    try:
        from learner.gnn_dagger import DAGGER                       # the reference, unmodified
        from learner.state_with_delay import MultiAgentStateWithDelay
        from learner.replay_buffer import Transition
    finally:
        sys.path.pop(0)
    assert os.path.realpath(sys.modules["learner.gnn_dagger"].__file__).startswith(os.path.realpath(REF))
    torch.manual_seed(seed)
    torch.set_num_threads(1)
    rng = np.random.RandomState(seed)
    device = torch.device("cpu")
    args = make_args(n_agents, k, hidden, n_layers, lr)
    learner = DAGGER(device, args)
    if checkpoint:
        learner.load_model(os.path.join(REF, checkpoint), "cpu")
    sd0 = {k_: v.detach().numpy().copy() for k_, v in learner.actor.state_dict().items()}

    # a pool of (state, expert label) pairs from an expert-driven rollout
    env = flock_env.FlockingRelativeOracle(n_agents=n_agents, comm_radius=comm_radius, v_max=v_max, rng=rng)
    env_state = env.reset()
    state = MultiAgentStateWithDelay(device, args, env_state, prev_state=None)
    states, labels = [], []
    for _ in range(pool):
        u = env.controller(centralized=False)
        label = torch.Tensor(u).transpose(1, 0).reshape((1, 1, 2, n_agents))        # gnn_dagger.py:173-175
        states.append(state)
        labels.append(label)
        env_state, _, _, _ = env.step(u)
        state = MultiAgentStateWithDelay(device, args, env_state, prev_state=state)

    zs, ys, losses = [], [], []
    grads1, sd1 = None, None
    for s in range(steps):
        idx = rng.choice(pool, size=batch, replace=False)
        bstates = tuple(states[i] for i in idx)
        blabels = tuple(labels[i] for i in idx)
        with torch.no_grad():
            z = torch.cat([torch.matmul(st.delay_state, st.delay_gso) for st in bstates])   # (B,K,F,N)
        zs.append(z.numpy().copy())
        ys.append(torch.cat(blabels).numpy().copy())
        loss = learner.gradient_step(Transition(bstates, blabels, None, None, None))
        losses.append(loss)
        if s == 0:
            grads1 = {n_: p.grad.detach().numpy().copy() for n_, p in learner.actor.named_parameters()}
            sd1 = {k_: v.detach().numpy().copy() for k_, v in learner.actor.state_dict().items()}
    sdT = {k_: v.detach().numpy().copy() for k_, v in learner.actor.state_dict().items()}
    out = dict(n_agents=n_agents, k=k, hidden=hidden, n_layers=n_layers, batch=batch, steps=steps, lr=lr, seed=seed,
               z=np.stack(zs), target=np.stack(ys), loss=np.array(losses, dtype=np.float64))
    for k_, v in sd0.items():
        out["sd0." + k_] = v
    for k_, v in sd1.items():
        out["sd1." + k_] = v
    for k_, v in sdT.items():
        out["sdT." + k_] = v
    for k_, v in grads1.items():
        out["grad1." + k_] = v
    os.makedirs(OUT, exist_ok=True)
    np.savez_compressed(os.path.join(OUT, name + ".npz"), **out)
    print(name, "N", n_agents, "K", k, "H", hidden, "L", n_layers, "B", batch, "steps", steps, "loss", losses[0], "->",
          losses[-1])


CASES = [
    dict(name="train_ckpt_n100_k3", n_agents=100, k=3, hidden=32, n_layers=2, batch=20, steps=4, lr=5e-5, seed=11,
         checkpoint=CKPT),
    dict(name="train_n50_k2_h16_l3", n_agents=50, k=2, hidden=16, n_layers=3, batch=7, steps=5, lr=1e-3, seed=7),
    dict(name="train_n64_k4_h64_l2", n_agents=64, k=4, hidden=64, n_layers=2, batch=5, steps=3, lr=1e-3, seed=3),
    dict(name="train_n12_k1_h4_l1", n_agents=12, k=1, hidden=4, n_layers=1, batch=3, steps=3, lr=1e-2, seed=5),
    dict(name="train_n33_k3_h128_l4", n_agents=33, k=3, hidden=128, n_layers=4, batch=4, steps=3, lr=1e-4, seed=13),
]


def main():
    if not os.path.isdir(REF):
        sys.exit("reference tree not found at " + REF)
    for c in CASES:
        run_case(**c)


if __name__ == "__main__":
    main()
