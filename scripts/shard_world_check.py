"""In-process sharded world on ONE GPU (the ranks are engine handles on their own streams, halo over p2p stores) against the
unsharded engine, bit for bit, at sizes where the warp-tiled adjacency and long grid rows are in play.
    python scripts/shard_world_check.py [world] [n_total] [steps]"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
from bench import make_weights      # noqa: E402
from oracle_free_state import synthetic_state  # noqa: E402
import test_gpu_sharding as tgs     # noqa: E402
from multiagent_gnn_policies_b200.engine import FlockEngine        # noqa: E402

world = int(sys.argv[1]) if len(sys.argv) > 1 else 8
n_total = int(sys.argv[2]) if len(sys.argv) > 2 else 160000
steps = int(sys.argv[3]) if len(sys.argv) > 3 else 12
sd, _ = make_weights(32, 3, 2)
x0 = synthetic_state(n_total, seed=31, density=1.6)
x0 = x0[np.argsort(x0[:, 0], kind="stable")]
single = FlockEngine(n_agents=n_total, k=3, hidden=32, n_layers=2, comm_radius=1.0, dt=0.01, edge_capacity=32)
single.load_state_dict(sd)
single.reset(x0)
import torch
from multiagent_gnn_policies_b200 import parallel
ranges = parallel.shard_ranges(n_total, world)
flocks = []
side = np.sqrt(n_total / 1.6)
for rank, (lo, cnt) in enumerate(ranges):
    stream = torch.cuda.Stream()
    be = parallel.CudaShardBackend(n_total, lo, cnt, ghost_capacity=n_total // 2, k=3, hidden=32, n_layers=2, comm_radius=1.0, dt=0.01,
                                   edge_capacity=32, stream=stream.cuda_stream, grid_dim=int(side / world + 40), grid_dim_y=int(side + 8))
    be._stream_keepalive = stream
    be.engine.load_state_dict(sd)
    flocks.append(parallel.ShardedFlock(be, rank, world, 3, 1.0, n_total // 4, all_gather=None))
got, bounds = tgs.drive(flocks, ranges, x0, steps, know_all=False, p2p=True)
bad = 0
for t in range(steps):
    single.step(None, None)
    if not np.array_equal(got[t], single.get_state()):
        bad += 1
        d = np.abs(got[t] - single.get_state()).max(axis=1)
        print("step", t, "differs at", int((d > 0).sum()), "agents, first", np.nonzero(d > 0)[0][:8])
print("world", world, "n", n_total, "steps", steps, "mismatching steps", bad, "overflow", [f.backend.overflow() for f in flocks],
      "stats0", flocks[0].backend.engine.stats())
