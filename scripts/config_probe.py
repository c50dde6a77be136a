"""Per-kernel CUDA-event times of one step for a few BASELINE configs (where does a config spend its step?).
    python scripts/config_probe.py"""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from bench import make_workload, make_weights      # noqa: E402
from multiagent_gnn_policies_b200.engine import FlockEngine        # noqa: E402

for name, n, b, hidden, radius in (("C2 N=10k H=64", 10_000, 1, 64, 1.0), ("C4 N=100k R=1", 100_000, 1, 32, 1.0),
                                   ("C4 N=100k R=2", 100_000, 1, 32, 2.0), ("C4 N=100k R=4", 100_000, 1, 32, 4.0),
                                   ("N=1M H=64", 1_000_000, 1, 64, 1.0)):
    sd, _ = make_weights(hidden, 3, 2)
    cap = int(max(24, 3.2 * np.pi * radius ** 2 * 1.6 + 16))
    eng = FlockEngine(n_agents=n, n_episodes=b, k=3, hidden=hidden, n_layers=2, comm_radius=radius, dt=0.01, edge_capacity=cap)
    eng.load_state_dict(sd)
    eng.reset(make_workload(n, seed=11))
    eng.rollout(20)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize()
    e0.record()
    eng.rollout(100)
    e1.record()
    torch.cuda.synchronize()
    per = {}
    for _ in range(5):
        for kname, ms in eng.profile_step():
            per[kname] = per.get(kname, 0.0) + ms / 5
    print(f"{name}: {e0.elapsed_time(e1) / 100 * 1e3:.1f} us/step deg {eng.stats()['n_edges'] / n:.1f} | "
          + " ".join(f"{k}={v * 1e3:.1f}" for k, v in per.items()), flush=True)
    eng.close()
