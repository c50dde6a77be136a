"""Sharded engine on one GPU: several engine handles act as ranks in one process (the all-gather is a
torch.stack) and must reproduce a single unsharded engine bit for bit, including across ownership hand-overs."""
import numpy as np
import pytest

from conftest import load_golden
from oracle import flock_env

pytestmark = pytest.mark.gpu


def make_world(n_total, world, sd, k=3, hidden=32, comm_radius=1.0, cap=None, own_streams=False, **kw):
    import torch
    from multiagent_gnn_policies_b200 import parallel
    ranges = parallel.shard_ranges(n_total, world)
    cap = cap or n_total
    flocks = []
    for rank, (lo, cnt) in enumerate(ranges):
        if own_streams:                          # p2p: a rank's step graph waits on the device for its peers' flags
            stream = torch.cuda.Stream()
            kw = dict(kw, stream=stream.cuda_stream)
        be = parallel.CudaShardBackend(n_total, lo, cnt, ghost_capacity=n_total, k=k, hidden=hidden,
                                       n_layers=2, comm_radius=comm_radius, dt=0.01, edge_capacity=64, **kw)
        if own_streams:
            be._stream_keepalive = stream
        be.engine.load_state_dict(sd)
        flocks.append(parallel.ShardedFlock(be, rank, world, k, comm_radius, cap, all_gather=None))
    return flocks, ranges


def drive(flocks, ranges, x0, steps, know_all=True, graphs=False, frame_velocity=0.0, p2p=False):
    """Lock-step driver: what ShardedFlock.reset/step do, with the collective replaced by a stack."""
    from multiagent_gnn_policies_b200 import parallel
    import torch
    world = len(flocks)
    bounds = parallel.strip_bounds(x0, ranges)
    stride = (flocks[0].cap + 1) * parallel.RECORD
    shared = torch.zeros((world, flocks[0].cap + 1, parallel.RECORD), dtype=torch.float64, device="cuda")

    def gather():
        torch.cuda.synchronize()                 # (the ranks may run on their own streams)
        shared.copy_(torch.stack([f.send for f in flocks]))
        torch.cuda.synchronize()
        for f in flocks:
            f.recv = shared

    for f in flocks:
        x_known = x0.copy()
        if not know_all:
            lo, cnt = ranges[f.rank]
            own_x = x0[lo:lo + cnt, 0]
            far = (x0[:, 0] < own_x.min() - f.depth) | (x0[:, 0] > own_x.max() + f.depth)
            far[lo:lo + cnt] = False
            x_known[far, 0] = parallel.FAR
        f.backend.configure(bounds, world, f.rank, f.depth, f.handover_margin, frame_velocity * f.dt, f.k + 1)
        f.backend.reset(x_known)
        win = np.zeros((world, parallel.RECORD))
        for q, (lo, cnt) in enumerate(ranges):
            win[q, 1], win[q, 2] = x0[lo:lo + cnt, 0].min(), x0[lo:lo + cnt, 0].max()
        f.windows0.copy_(torch.from_numpy(win))
        f.backend.pack(f.windows0.reshape(-1)[1:], parallel.RECORD, f.send, f.cap, False)
    gather()
    for f in flocks:
        f.backend.unpack(shared, f.cap)
        f.backend.build(False)
    if p2p:
        torch.cuda.synchronize()
        parallel.connect_p2p_local(flocks)
        torch.cuda.synchronize()
    out = []
    for _ in range(steps):
        if p2p:                                  # one graph per rank, launched on the ranks' own streams; the halo records
            for f in flocks:                     # travel as plain stores into the peers' inboxes, flags order them
                f.backend.step_p2p()
            torch.cuda.synchronize()
        elif graphs:                             # CUDA-graph replayed halves
            for f in flocks:
                f.backend.step_begin(shared.reshape(-1)[1:], stride, f.send, f.cap)
            gather()
            for f in flocks:
                f.backend.step_end(shared, f.cap)
        else:
            for f in flocks:
                f.backend.local_step()
                f.backend.pack(shared.reshape(-1)[1:], stride, f.send, f.cap, True)
            gather()
            for f in flocks:
                f.backend.unpack(shared, f.cap)
                f.backend.build(True)
        n_total = x0.shape[0]
        owners = np.zeros(n_total, int)
        x_all = np.zeros((n_total, 4))
        for f in flocks:
            ids, st = f.backend.owned_state()
            owners[ids] += 1
            x_all[ids] = st
        assert np.all(owners == 1)               # exactly one owner per agent at all times
        out.append(x_all)
    return out, bounds


@pytest.mark.parametrize("world,n_total,order,graphs", [(2, 3000, "sorted", False), (4, 5000, "sorted", True),
                                                        (3, 1500, "random", False), (2, 2000, "random", True),
                                                        (2, 3000, "sorted", "p2p"), (4, 5000, "sorted", "p2p"),
                                                        (3, 1500, "random", "p2p")])
def test_sharded_equals_single_engine(world, n_total, order, graphs):
    from multiagent_gnn_policies_b200.engine import FlockEngine
    g = load_golden("ckpt_n100_k3")
    x0 = flock_env.synthetic_state(n_total, seed=31, density=1.6)
    if order == "sorted":
        x0 = x0[np.argsort(x0[:, 0], kind="stable")]
    else:
        x0 = x0[np.random.RandomState(1).permutation(n_total)]
    steps = 60 if order == "sorted" else 10
    single = FlockEngine(n_agents=n_total, k=3, hidden=32, n_layers=2, comm_radius=1.0, dt=0.01, edge_capacity=64)
    single.load_state_dict(g["state_dict"])
    single.reset(x0)
    p2p = graphs == "p2p"
    flocks, ranges = make_world(n_total, world, g["state_dict"], own_streams=p2p)
    got, bounds = drive(flocks, ranges, x0, steps, know_all=(order != "sorted"), graphs=bool(graphs) and not p2p, p2p=p2p)
    for t in range(steps):
        single.step(None, None)
        # same kernels, same per-row neighbour order (cell lists are canonical) -> identical bits,
        # whoever owns an agent and however often it changed hands
        np.testing.assert_array_equal(got[t], single.get_state())
    moved = 0
    for f in flocks:
        lo, cnt = ranges[f.rank]
        moved += int(not np.array_equal(np.sort(f.backend.owned()), np.arange(lo, lo + cnt)))
    if order == "random":
        assert moved > 0                          # ownership was handed over (re-partitioning)
        x_end = single.get_state()
        for f in flocks:                          # ownership re-partitioned itself into the x-strips
            xs = x_end[f.backend.owned(), 0]
            assert xs.min() >= bounds[f.rank] - 1.6 and xs.max() <= bounds[f.rank + 1] + 1.6
    for f in flocks:
        assert not f.backend.overflow()
        f.backend.engine.close()
    single.close()
