"""Generate tests/golden/agg_*.npz: the UNMODIFIED reference ``Actor.forward`` (learner/actor.py:45-86) for aggregation
indices other than DAGGER's 0 and for layer widths the rollout engine does not template (unequal hidden widths).

TEST INFRASTRUCTURE.  Run in the build container only (needs /root/reference):

    python -m oracle.gen_golden_agg

Inputs are real state containers -- the reference's own ``MultiAgentStateWithDelay`` (learner/state_with_delay.py:6-53)
over a few steps of the spec env (oracle.flock_env), B consecutive states concatenated the way ``gradient_step``
batches them (learner/gnn_dagger.py:84-86) -- so the delayed operators are genuine products A_t ... A_{t-k+1}.
These fixtures pin ``oracle.learner.actor_forward_any`` and the CUDA path ``fgnn_actor_forward_general``.
"""
import os
import sys

import numpy as np
import torch

from oracle import flock_env
from oracle.gen_golden import OUT, REF, _load, make_args

CASES = [
    # name, N, K, layer widths (n_s .. n_a), ind_agg, batch, seed
    dict(name="agg1_n40_k3", n_agents=40, k=3, layers=[6, 16, 24, 2], ind_agg=1, batch=2, seed=21),
    dict(name="agg2_n30_k2", n_agents=30, k=2, layers=[6, 8, 8, 8, 2], ind_agg=2, batch=1, seed=22),
    dict(name="agglast_n20_k4", n_agents=20, k=4, layers=[6, 12, 2], ind_agg=1, batch=3, seed=23),
    dict(name="agg0_uneven_n25_k3", n_agents=25, k=3, layers=[6, 20, 10, 2], ind_agg=0, batch=2, seed=24),
    dict(name="agg1_n150_k3_h128", n_agents=150, k=3, layers=[6, 128, 128, 2], ind_agg=1, batch=2, seed=25),
]


def run_case(name, n_agents, k, layers, ind_agg, batch, seed):
    ref_actor = _load("ref_actor", "learner/actor.py")
    ref_state = _load("ref_state", "learner/state_with_delay.py")
    torch.manual_seed(seed)
    torch.set_num_threads(1)
    actor = ref_actor.Actor(layers[0], layers[-1], list(layers[1:-1]), k, ind_agg)
    actor.eval()
    env = flock_env.FlockingRelativeOracle(n_agents=n_agents, comm_radius=1.0, v_max=3.0, dt=0.01,
                                           rng=np.random.RandomState(seed))
    env_state = env.reset()
    args = make_args(n_agents, k)
    device = torch.device("cpu")
    state = ref_state.MultiAgentStateWithDelay(device, args, env_state, prev_state=None)
    states = []
    for t in range(k + batch):
        u = env.controller()
        env_state, _, _, _ = env.step(u)
        state = ref_state.MultiAgentStateWithDelay(device, args, env_state, prev_state=state)
        if t >= k:
            states.append(state)
    ds = torch.cat([s.delay_state for s in states])                      # gnn_dagger.py:84-85
    gso = torch.cat([s.delay_gso for s in states])
    with torch.no_grad():
        out = actor(ds, gso)
    res = dict(n_agents=n_agents, k=k, layers=np.array(layers), ind_agg=ind_agg, batch=batch, seed=seed,
               delay_state=ds.numpy(), delay_gso=gso.numpy(), out=out.numpy())
    for k_, v in actor.state_dict().items():
        res["sd." + k_] = v.detach().numpy()
    np.savez_compressed(os.path.join(OUT, name + ".npz"), **res)
    print(name, "N", n_agents, "K", k, "layers", layers, "ind_agg", ind_agg, "B", batch, "|out|max", float(out.abs().max()))


def main():
    if not os.path.isdir(REF):
        sys.exit("reference tree not found at " + REF)
    for c in CASES:
        run_case(**c)


if __name__ == "__main__":
    main()
