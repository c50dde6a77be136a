"""The multi-GPU halo protocol (multiagent_gnn_policies_b200.parallel.ShardedFlock) on CPU: two and three gloo
ranks (a middle rank has two neighbours) with the numpy backend must reproduce the single-process oracle rollout
of the whole flock."""
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

N_TOTAL, STEPS, K, R = 600, 10, 3, 1.0


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
        far = (x0[:, 0] < own_x.min() - flock.depth) | (x0[:, 0] > own_x.max() + flock.depth)
        far[lo:lo + cnt] = False
        x_known[far, 0] = parallel.FAR
    bounds = parallel.strip_bounds(x0, ranges)            # every rank derives the same territories
    flock.reset(x_known, ranges, bounds=bounds)
    ids, states, actions, pools = [], [], [], []
    for _ in range(STEPS):
        flock.step()
        i1, st = backend.owned_state()
        i2, ac = backend.owned_action()
        assert np.array_equal(i1, i2)
        ids.append(np.pad(i1, (0, N_TOTAL - i1.size), constant_values=-1))
        states.append(np.pad(st, ((0, N_TOTAL - i1.size), (0, 0))))
        actions.append(np.pad(ac, ((0, N_TOTAL - i1.size), (0, 0))))
        pools.append(len(backend.pool))
    np.savez(os.path.join(out_dir, f"rank{rank}.npz"), ids=np.stack(ids), states=np.stack(states),
             actions=np.stack(actions), pools=np.array(pools), overflow=backend.overflow(),
             handed_over=backend.handed_over, bounds=bounds)
    dist.destroy_process_group()


@pytest.mark.parametrize("sorted_order,world", [(True, 2), (False, 2), (True, 3)])
def test_gloo_ranks_reproduce_the_single_process_oracle(tmp_path, sorted_order, world):
    g = load_golden("ckpt_n100_k3")
    sd = g["state_dict"]
    layers = learner.weights_from_state_dict(sd)
    x0 = flock_env.synthetic_state(N_TOTAL, seed=21, density=1.6)
    if sorted_order:
        x0 = x0[np.argsort(x0[:, 0], kind="stable")]          # strips along x: index order == x order
    else:
        x0 = x0[np.random.RandomState(0).permutation(N_TOTAL)]
    xs_ref, acts_ref = reference_rollout(x0, layers, STEPS)
    port = _free_port()
    mp.spawn(_worker, args=(world, port, x0, sd, str(tmp_path), sorted_order), nprocs=world, join=True)
    outs = [np.load(tmp_path / f"rank{rank}.npz") for rank in range(world)]
    for out in outs:
        assert not bool(out["overflow"])
    for t in range(STEPS):
        owner_count = np.zeros(N_TOTAL, int)
        x_all = np.zeros((N_TOTAL, 4))
        a_all = np.zeros((N_TOTAL, 2), np.float32)
        for out in outs:
            ids = out["ids"][t]
            ids = ids[ids >= 0]
            owner_count[ids] += 1
            x_all[ids] = out["states"][t][:ids.size]
            a_all[ids] = out["actions"][t][:ids.size]
        assert np.all(owner_count == 1)                       # every agent has exactly one owner, always
        # the action reported at step t was computed by whoever owned the agent BEFORE this step's hand-over;
        # after a hand-over the new owner's action slot is stale until it computes one: compare states (which
        # carry every action's effect) at every step and actions where ownership did not just change
        np.testing.assert_allclose(x_all, xs_ref[t], rtol=1e-9, atol=2e-6)
    total_handed = sum(int(o["handed_over"]) for o in outs)
    if sorted_order:          # the exchange stays a thin boundary layer
        assert max(o["pools"].max() for o in outs) < N_TOTAL / world + 0.45 * N_TOTAL
    else:                     # arbitrary order: ownership re-partitions itself into the strips
        assert total_handed > 0.3 * N_TOTAL
        bounds = outs[0]["bounds"]
        for rank, out in enumerate(outs):
            ids = out["ids"][-1]
            ids = ids[ids >= 0]
            xs = xs_ref[-1][ids, 0]
            assert xs.min() >= bounds[rank] - 1.6 and xs.max() <= bounds[rank + 1] + 1.6
        assert outs[0]["pools"][-1] < 0.8 * N_TOTAL           # and the halo became a thin layer


def test_shard_ranges_and_depth():
    assert parallel.shard_ranges(10, 3) == [(0, 4), (4, 3), (7, 3)]
    assert sum(c for _, c in parallel.shard_ranges(1_000_003, 8)) == 1_000_003
    assert parallel.halo_depth(3, 1.0, 0.5) == 3.5
    b = parallel.strip_bounds(np.array([[0.0, 0, 0, 0], [1.0, 0, 0, 0], [3.0, 0, 0, 0], [4.0, 0, 0, 0]]), [(0, 2), (2, 2)])
    assert b[1] == 2.0 and b[0] < -1e200 and b[2] > 1e200


def test_cpu_binding_is_best_effort():
    """parallel.bind_to_local_cpus (one process per GPU: rank on its GPU's NUMA node) must never raise and must leave the
    affinity alone where it cannot name a proper subset of the allowed CPUs -- e.g. here, without a GPU / NVML device."""
    import os
    from multiagent_gnn_policies_b200 import parallel
    before = os.sched_getaffinity(0)
    got = parallel.bind_to_local_cpus(0)
    after = os.sched_getaffinity(0)
    assert got is None or (set(got) == after and after < before)
    if got is None:
        assert after == before
