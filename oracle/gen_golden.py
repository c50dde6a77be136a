"""Generate tests/golden/*.npz by running the UNMODIFIED reference learner code.

TEST INFRASTRUCTURE.  Run in the build container only (needs /root/reference):

    python -m oracle.gen_golden

The reference modules are loaded by file path under alias names
(learner/actor.py, learner/state_with_delay.py) so they never shadow anything of ours.
The env side is oracle.flock_env (parity unpinned, see its header); the learner side --
MultiAgentStateWithDelay + Actor.forward + the select_action reshaping
(learner/gnn_dagger.py:63-70) -- is the reference itself, so these files pin
oracle.learner and the CUDA path.
"""
import configparser
import importlib.util
import os
import sys

import numpy as np
import torch

from oracle import flock_env

REF = os.environ.get("FGNN_REFERENCE", "/root/reference")
OUT = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden")


def _load(alias, rel):
    spec = importlib.util.spec_from_file_location(alias, os.path.join(REF, rel))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


def make_args(n_agents, k, n_states=6):
    cp = configparser.ConfigParser()
    cp.read_dict({"DEFAULT": {"n_states": str(n_states), "n_agents": str(n_agents), "k": str(k)}})
    return cp["DEFAULT"]


def run_case(name, n_agents, k, hidden, n_layers, steps, seed, checkpoint=None,
             comm_radius=1.0, dt=0.01, v_max=3.0, synthetic_density=None):
    ref_actor = _load("ref_actor", "learner/actor.py")
    ref_state = _load("ref_state", "learner/state_with_delay.py")
    torch.manual_seed(seed)
    torch.set_num_threads(1)
    rng = np.random.RandomState(seed)
    actor = ref_actor.Actor(6, 2, [hidden] * n_layers, k, 0)
    if checkpoint:
        actor.load_state_dict(torch.load(os.path.join(REF, checkpoint), map_location="cpu"))
    actor.eval()
    env = flock_env.FlockingRelativeOracle(n_agents=n_agents, comm_radius=comm_radius, v_max=v_max,
                                           dt=dt, rng=rng)
    if synthetic_density is None:
        env_state = env.reset()
    else:
        env_state = env.reset(flock_env.synthetic_state(n_agents, seed=seed, density=synthetic_density,
                                                        v_max=v_max))
    args = make_args(n_agents, k)
    device = torch.device("cpu")
    state = ref_state.MultiAgentStateWithDelay(device, args, env_state, prev_state=None)
    xs, values, degs, zs, actions, rewards = [], [], [], [], [], []
    for t in range(steps):
        xs.append(env.x.copy())
        values.append(env_state[0].copy())
        degs.append((env_state[1] != 0).sum(axis=1).astype(np.int32))
        with torch.no_grad():
            z = torch.matmul(state.delay_state, state.delay_gso)             # actor.py:70 (B,K,F,N)
            mu = actor(state.delay_state, state.delay_gso)                    # (1,1,2,N)
            mu = mu.permute(0, 1, 3, 2).reshape(n_agents, 2)                  # gnn_dagger.py:67-68
        zs.append(z[0].numpy().copy())
        actions.append(mu.numpy().copy())
        env_state, reward, _, _ = env.step(mu.numpy())
        rewards.append(reward)
        state = ref_state.MultiAgentStateWithDelay(device, args, env_state, prev_state=state)
    sd = {k_: v.detach().numpy() for k_, v in actor.state_dict().items()}
    out = dict(
        n_agents=n_agents, k=k, hidden=hidden, n_layers=n_layers, steps=steps, seed=seed,
        comm_radius=comm_radius, dt=dt, v_max=v_max,
        x=np.stack(xs), values=np.stack(values), deg=np.stack(degs), z=np.stack(zs),
        action=np.stack(actions), reward=np.array(rewards),
    )
    for k_, v in sd.items():
        out["sd." + k_] = v
    os.makedirs(OUT, exist_ok=True)
    np.savez_compressed(os.path.join(OUT, name + ".npz"), **out)
    print(name, "N", n_agents, "K", k, "H", hidden, "L", n_layers, "T", steps,
          "|a|max", float(np.abs(out["action"]).max()), "deg mean", float(out["deg"].mean()))


CKPT = "models/actor_FlockingRelative-v0_dagger_k3"

CASES = [
    # name, N, K, H, L, steps, seed, checkpoint
    dict(name="ckpt_n100_k3", n_agents=100, k=3, hidden=32, n_layers=2, steps=8, seed=11, checkpoint=CKPT),
    dict(name="ckpt_n400_k3_uniform", n_agents=400, k=3, hidden=32, n_layers=2, steps=5, seed=3,
         checkpoint=CKPT, synthetic_density=1.6),
    dict(name="rand_n12_k1_h4_l1", n_agents=12, k=1, hidden=4, n_layers=1, steps=3, seed=5),
    dict(name="rand_n50_k2_h16_l3", n_agents=50, k=2, hidden=16, n_layers=3, steps=5, seed=7),
    dict(name="rand_n100_k4_h64_l2", n_agents=100, k=4, hidden=64, n_layers=2, steps=7, seed=11),
    dict(name="rand_n64_k3_h128_l4", n_agents=64, k=3, hidden=128, n_layers=4, steps=5, seed=13),
    dict(name="rand_n300_k3_h64_l2_r2", n_agents=300, k=3, hidden=64, n_layers=2, steps=4, seed=17,
         comm_radius=2.0, synthetic_density=1.6),
]


def main():
    if not os.path.isdir(REF):
        sys.exit("reference tree not found at " + REF)
    for c in CASES:
        run_case(**c)


if __name__ == "__main__":
    main()
