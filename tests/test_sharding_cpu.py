"""The multi-GPU halo protocol (multiagent_gnn_policies_b200.parallel.ShardedFlock) on CPU: two gloo ranks
with the numpy backend must reproduce the single-process oracle rollout of the whole flock."""
import os
import socket
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from conftest import ROOT, load_golden
from oracle import flock_env, learner, sparse
from multiagent_gnn_policies_b200 import parallel

N_TOTAL, STEPS, K, R = 600, 6, 3, 1.0


def reference_rollout(x0, layers, steps):
    """Single-process oracle: closed loop over the whole flock; returns states and actions per step."""
    x = x0.copy()
    sstate, xs, acts = None, [], []
    for _ in range(steps):
        sv, deg, i, j = sparse.compute_helpers_sparse(x, R)
        sstate = sparse.SparseDelayState(sv, sparse.network_csr(x.shape[0], deg, i, j), prev_state=sstate, k=K)
        a = sparse.readout(layers, sstate.aggregate())
        x = flock_env.integrate(x, a, 0.01)
        xs.append(x.copy())
        acts.append(a)
    return xs, acts


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, x0, sd, out_dir, sorted_order):
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    from shard_numpy_backend import NumpyShardBackend
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    layers = learner.weights_from_state_dict(sd)
    ranges = parallel.shard_ranges(N_TOTAL, world)
    lo, cnt = ranges[rank]
    cap = N_TOTAL            # generous: the random-order case sends everything
    backend = NumpyShardBackend(N_TOTAL, lo, cnt, layers, k=K, comm_radius=R)

    def all_gather(send):
        t = torch.from_numpy(np.ascontiguousarray(send))
        outs = [torch.zeros_like(t) for _ in range(world)]
        dist.all_gather(outs, t)
        return torch.stack(outs).numpy()

    flock = parallel.ShardedFlock(backend, rank, world, K, R, cap, all_gather)
    # each rank only knows the agents near its own strip at reset; the rest sits at FAR
    x_known = x0.copy()
    if sorted_order:
        own_x = x0[lo:lo + cnt, 0]
        far = (x0[:, 0] < own_x.min() - flock.send_depth) | (x0[:, 0] > own_x.max() + flock.send_depth)
        far[lo:lo + cnt] = False
        x_known[far, 0] = parallel.FAR
    flock.reset(x_known, ranges)
    states, actions, pools = [], [], []
    for _ in range(STEPS):
        flock.step()
        states.append(backend.owned_state().copy())
        actions.append(backend.owned_action().copy())
        pools.append(len(backend.pool))
    np.savez(os.path.join(out_dir, f"rank{rank}.npz"), states=np.stack(states), actions=np.stack(actions),
             pools=np.array(pools), overflow=backend.overflow())
    dist.destroy_process_group()


@pytest.mark.parametrize("sorted_order", [True, False])
def test_two_gloo_ranks_reproduce_the_single_process_oracle(tmp_path, sorted_order):
    g = load_golden("ckpt_n100_k3")
    sd = g["state_dict"]
    layers = learner.weights_from_state_dict(sd)
    x0 = flock_env.synthetic_state(N_TOTAL, seed=21, density=1.6)
    if sorted_order:
        x0 = x0[np.argsort(x0[:, 0], kind="stable")]          # strips along x: index order == x order
    else:
        x0 = x0[np.random.RandomState(0).permutation(N_TOTAL)]
    xs_ref, acts_ref = reference_rollout(x0, layers, STEPS)
    world = 2
    port = _free_port()
    mp.spawn(_worker, args=(world, port, x0, sd, str(tmp_path), sorted_order), nprocs=world, join=True)
    ranges = parallel.shard_ranges(N_TOTAL, world)
    for rank, (lo, cnt) in enumerate(ranges):
        out = np.load(tmp_path / f"rank{rank}.npz")
        assert not bool(out["overflow"])
        for t in range(STEPS):
            np.testing.assert_allclose(out["actions"][t], acts_ref[t][lo:lo + cnt], rtol=2e-5, atol=2e-5)
            np.testing.assert_allclose(out["states"][t], xs_ref[t][lo:lo + cnt], rtol=1e-9, atol=1e-7)
        if sorted_order:      # the exchange stays a thin boundary layer
            assert out["pools"].max() < cnt + 0.45 * N_TOTAL
        else:                 # arbitrary order: windows overlap, everything is present everywhere
            assert out["pools"].max() == N_TOTAL


def test_shard_ranges_and_depth():
    assert parallel.shard_ranges(10, 3) == [(0, 4), (4, 3), (7, 3)]
    assert sum(c for _, c in parallel.shard_ranges(1_000_003, 8)) == 1_000_003
    assert parallel.halo_depth(3, 1.0, 0.5) == 3.5
