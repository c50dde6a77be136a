"""Workload for compute-sanitizer (memcheck / racecheck / synccheck): a 10k-agent flock through every kernel of the step --
k_adjacency_t (FGNN_STEP_MODE=0), the warp-tiled k_pair_adjacency with its TMA staging (FGNN_STEP_MODE=1), graph replay, the API-split path with
float32 and float64 actions, the expert controller, the single-CTA kernels of small flocks, and two in-process ranks
over the p2p halo transport (folded prepare / flag).
    compute-sanitizer --tool memcheck python scripts/sanitize_step.py"""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from bench import make_workload, make_weights                      # noqa: E402
from multiagent_gnn_policies_b200 import parallel                  # noqa: E402
from multiagent_gnn_policies_b200.engine import FlockEngine        # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 10_000
sd, _ = make_weights(32, 3, 2)
x0 = make_workload(n)
for mode in ("0", "1"):
    os.environ["FGNN_STEP_MODE"] = mode
    eng = FlockEngine(n_agents=n, k=3, hidden=32, n_layers=2, comm_radius=1.0, dt=0.01)
    eng.load_state_dict(sd)
    eng.reset(x0)
    for _ in range(4):
        eng.step(None, None)
    eng.rollout(4)
    a = eng.policy().cpu().numpy()
    eng.env_step(a)
    u = eng.controller(centralized=False, dtype=np.float64)
    eng.env_step(u)
    eng.controller(centralized=True)
    eng.sync()
    print("mode", mode, eng.stats())
    eng.close()
# the single-CTA kernels of small flocks (fgnn_mini.cu): fused step, rollout, policy, env_step (fp32 and float64 actions)
for nb, ne in ((100, 1), (30, 4)):
    eng = FlockEngine(n_agents=nb, n_episodes=ne, k=3, hidden=32, n_layers=2, comm_radius=1.0, dt=0.01)
    eng.load_state_dict(sd)
    eng.reset(np.concatenate([make_workload(nb, seed=3 + e) for e in range(ne)]))
    for _ in range(3):
        eng.step(None, None)
    eng.rollout(5)
    a = eng.policy().cpu().numpy()
    eng.env_step(a)
    eng.env_step(a.astype(np.float64))
    eng.host_rows = eng.get_features()
    eng.sync()
    print("mini", nb, ne, eng.stats())
    eng.close()
os.environ["FGNN_STEP_MODE"] = "0"
x0 = x0[np.argsort(x0[:, 0], kind="stable")]
ranges = parallel.shard_ranges(n, 2)
flocks, streams = [], []
for rank, (lo, cnt) in enumerate(ranges):
    streams.append(torch.cuda.Stream())
    be = parallel.CudaShardBackend(n, lo, cnt, ghost_capacity=n, k=3, hidden=32, n_layers=2, comm_radius=1.0, dt=0.01,
                                   edge_capacity=64, stream=streams[-1].cuda_stream)
    be.engine.load_state_dict(sd)
    flocks.append(parallel.ShardedFlock(be, rank, 2, 3, 1.0, n, all_gather=None))
bounds = parallel.strip_bounds(x0, ranges)
shared = torch.zeros((2, n + 1, parallel.RECORD), dtype=torch.float64, device="cuda")
for f in flocks:
    f.backend.configure(bounds, 2, f.rank, f.depth, f.handover_margin, 0.0, f.k + 1)
    f.backend.reset(x0)
    win = np.zeros((2, parallel.RECORD))
    for q, (lo, cnt) in enumerate(ranges):
        win[q, 1], win[q, 2] = x0[lo:lo + cnt, 0].min(), x0[lo:lo + cnt, 0].max()
    f.windows0.copy_(torch.from_numpy(win))
    f.backend.pack(f.windows0.reshape(-1)[1:], parallel.RECORD, f.send, f.cap, False)
torch.cuda.synchronize()
shared.copy_(torch.stack([f.send for f in flocks]))
torch.cuda.synchronize()
for f in flocks:
    f.recv = shared
    f.backend.unpack(shared, f.cap)
    f.backend.build(False)
torch.cuda.synchronize()
parallel.connect_p2p_local(flocks)
torch.cuda.synchronize()
for _ in range(6):
    for f in flocks:
        f.backend.step_p2p()
    torch.cuda.synchronize()
print("p2p ranks", [f.backend.engine.stats() for f in flocks])
