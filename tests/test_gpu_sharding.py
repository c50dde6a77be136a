"""Sharded engine on one GPU: several engine handles act as ranks in one process (the all-gather is a
torch.cat), and must reproduce a single unsharded engine bit for bit on the owned agents."""
import numpy as np
import pytest

from conftest import load_golden
from oracle import flock_env

pytestmark = pytest.mark.gpu


class InProcessWorld:
    """Lock-step driver for `world` ShardedFlock objects living in one process."""

    def __init__(self, flocks):
        self.flocks = flocks
        self.pending = None

    def run_phase(self, fn_before, fn_after):
        import torch
        for f in self.flocks:
            fn_before(f)
        recv = torch.stack([f.send for f in self.flocks]).contiguous()
        for f in self.flocks:
            f.recv = recv
            fn_after(f)


def make_world(n_total, world, x0, sd, k=3, hidden=32, comm_radius=1.0, cap=None, **kw):
    from multiagent_gnn_policies_b200 import parallel
    ranges = parallel.shard_ranges(n_total, world)
    cap = cap or n_total
    flocks = []
    for rank, (lo, cnt) in enumerate(ranges):
        be = parallel.CudaShardBackend(n_total, lo, cnt, ghost_capacity=min(n_total, world * cap), k=k, hidden=hidden,
                                       n_layers=2, comm_radius=comm_radius, dt=0.01, edge_capacity=64, **kw)
        be.engine.load_state_dict(sd)
        flocks.append(parallel.ShardedFlock(be, rank, world, k, comm_radius, cap, all_gather=None))
    return flocks, ranges


def drive(flocks, ranges, x0, steps, know_all=True, graphs=False):
    from multiagent_gnn_policies_b200 import parallel
    import torch
    world = len(flocks)

    def exchange(windows_of, stride):
        for f in flocks:
            w = windows_of(f)
            f.backend.pack(w.reshape(-1)[1:], stride, world, f.rank, f.send_depth, f.send, f.cap)
        recv = torch.stack([f.send for f in flocks]).contiguous()
        for f in flocks:
            f.recv = recv
            f.backend.unpack(recv, world, f.rank, f.cap, f.depth)

    for f in flocks:
        x_known = x0.copy()
        if not know_all:
            lo, cnt = ranges[f.rank]
            own_x = x0[lo:lo + cnt, 0]
            far = (x0[:, 0] < own_x.min() - f.send_depth) | (x0[:, 0] > own_x.max() + f.send_depth)
            far[lo:lo + cnt] = False
            x_known[far, 0] = parallel.FAR
        f.backend.reset(x_known)
        win = np.zeros((world, parallel.RECORD))
        for q, (lo, cnt) in enumerate(ranges):
            win[q, 1], win[q, 2] = x0[lo:lo + cnt, 0].min(), x0[lo:lo + cnt, 0].max()
        f.windows0.copy_(torch.from_numpy(win))
    exchange(lambda f: f.windows0, parallel.RECORD)
    for f in flocks:
        f.backend.build(False)
    out = []
    stride = (flocks[0].cap + 1) * parallel.RECORD
    shared = flocks[0].recv.clone()              # persistent gathered buffer: stable pointers for the graph cache
    for f in flocks:
        f.recv = shared
    for _ in range(steps):
        if graphs:                               # CUDA-graph replayed halves
            for f in flocks:
                f.backend.step_begin(shared.reshape(-1)[1:], stride, world, f.rank, f.send_depth, f.send, f.cap)
            gathered = torch.stack([f.send for f in flocks])
            shared.copy_(gathered)
            for f in flocks:
                f.backend.step_end(shared, world, f.rank, f.cap, f.depth)
        else:
            for f in flocks:
                f.backend.local_step()
            exchange(lambda f: f.recv, stride)
            for f in flocks:
                f.backend.build(True)
            shared.copy_(flocks[0].recv)
            for f in flocks:
                f.recv = shared
        out.append((np.concatenate([f.backend.owned_state() for f in flocks]),
                    np.concatenate([f.backend.owned_action() for f in flocks])))
    return out


@pytest.mark.parametrize("world,n_total,order,graphs", [(2, 3000, "sorted", False), (4, 5000, "sorted", True),
                                                        (3, 1500, "random", False), (2, 2000, "random", True)])
def test_sharded_equals_single_engine(world, n_total, order, graphs):
    from multiagent_gnn_policies_b200.engine import FlockEngine
    g = load_golden("ckpt_n100_k3")
    x0 = flock_env.synthetic_state(n_total, seed=31, density=1.6)
    if order == "sorted":
        x0 = x0[np.argsort(x0[:, 0], kind="stable")]
    else:
        x0 = x0[np.random.RandomState(1).permutation(n_total)]
    steps = 8
    single = FlockEngine(n_agents=n_total, k=3, hidden=32, n_layers=2, comm_radius=1.0, dt=0.01, edge_capacity=64)
    single.load_state_dict(g["state_dict"])
    single.reset(x0)
    flocks, ranges = make_world(n_total, world, x0, g["state_dict"])
    got = drive(flocks, ranges, x0, steps, know_all=(order != "sorted"), graphs=graphs)
    for t in range(steps):
        a = np.empty((n_total, 2), np.float32)
        single.step(a, None)
        # same kernels, same per-row neighbour order (cell lists are canonical) -> identical bits
        np.testing.assert_array_equal(got[t][1], a)
        np.testing.assert_array_equal(got[t][0], single.get_state())
    for f in flocks:
        assert not f.backend.overflow()
        f.backend.engine.close()
    if order == "sorted":
        # thin boundary layer: a rank holds far fewer agents than the whole flock
        pool = flocks[0].backend.engine
        assert ranges[0][1] < n_total
    single.close()
