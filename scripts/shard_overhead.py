"""Where does a sharded rank's step spend more time than the unsharded engine?  One GPU, world = 1: the rank owns every agent
(no ghosts, nothing to exchange), so what is measured is the machinery itself -- owned / ghost list indirection, the fused
pack epilogue, the extra small kernels of the p2p step -- kernel by kernel (CUDA events), next to the plain engine.
    python scripts/shard_overhead.py [N] [steps]"""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from bench import make_workload, make_weights, DENSITY      # noqa: E402
from multiagent_gnn_policies_b200 import parallel                  # noqa: E402
from multiagent_gnn_policies_b200.engine import FlockEngine        # noqa: E402


def timed(fn, steps):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    best = 1e9
    for _ in range(3):
        torch.cuda.synchronize()
        e0.record()
        for _ in range(steps):
            fn()
        e1.record()
        torch.cuda.synchronize()
        best = min(best, e0.elapsed_time(e1) / steps)
    return best


def profile(eng, reps=5):
    per = {}
    for _ in range(reps):
        for name, ms in eng.profile_step():
            per[name] = per.get(name, 0.0) + ms / reps
    return per


def main():
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 1_000_000
    steps = int(sys.argv[2]) if len(sys.argv) > 2 else 200
    x0 = make_workload(n)
    sd, _ = make_weights(32, 3, 2)
    side = np.sqrt(n / DENSITY)
    cap = int(max(24, 3.2 * np.pi * DENSITY + 16))
    ms_plain, per_plain = 0.0, {}
    if os.environ.get("FGNN_SHARD_ONLY") != "1":
        eng = FlockEngine(n_agents=n, k=3, hidden=32, n_layers=2, comm_radius=1.0, dt=0.01, edge_capacity=cap)
        eng.load_state_dict(sd)
        eng.reset(x0)
        eng.rollout(30)
        ms_plain = timed(lambda: eng.rollout(1), steps)
        per_plain = profile(eng)
        eng.close()

    depth = parallel.halo_depth(3, 1.0)
    halo_cap = int(1.5 * (depth + 2.0) * side * DENSITY * 2) + 1024
    gx, gy = int(np.ceil(side + 2 * depth + 4)) + 2, int(np.ceil(side)) + 4
    be = parallel.CudaShardBackend(n, 0, n, ghost_capacity=2 * halo_cap, k=3, hidden=32, n_layers=2, comm_radius=1.0, dt=0.01,
                                   edge_capacity=cap, grid_dim=gx, grid_dim_y=gy)
    be.engine.load_state_dict(sd)
    flock = parallel.ShardedFlock(be, 0, 1, 3, 1.0, halo_cap, all_gather=lambda send: send.reshape(1, *send.shape))
    flock.reset(x0, [(0, n)], bounds=np.array([-parallel.INF, parallel.INF]))
    parallel.connect_p2p_local([flock])
    for _ in range(30):
        flock.step()
    ms_shard = timed(flock.step, steps)
    per_shard = profile(be.engine)
    fmt = lambda per: " ".join(f"{k}={v * 1e3:.1f}" for k, v in per.items())
    print(f"N={n}: plain engine {ms_plain * 1e3:.1f} us/step  [{fmt(per_plain)}] sum={sum(per_plain.values()) * 1e3:.1f}")
    print(f"N={n}: sharded rank, world=1, p2p step {ms_shard * 1e3:.1f} us/step  [{fmt(per_shard)}] sum={sum(per_shard.values()) * 1e3:.1f}")


if __name__ == "__main__":
    main()
