"""Config C3 (256 episodes x 1000 agents) kernel by kernel, for both adjacency kernels.
    python scripts/c3_probe.py"""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from bench import make_workload, make_weights      # noqa: E402
from multiagent_gnn_policies_b200.engine import FlockEngine        # noqa: E402

sd, _ = make_weights(32, 3, 2)
n, b = 1000, 256
xs = np.concatenate([make_workload(n, seed=11 + e) for e in range(4)])
xs = np.concatenate([xs] * (b // 4))
for mode in ("0", "1"):
    os.environ["FGNN_STEP_MODE"] = mode
    eng = FlockEngine(n_agents=n, n_episodes=b, k=3, hidden=32, n_layers=2, comm_radius=1.0, dt=0.01, edge_capacity=32)
    eng.load_state_dict(sd)
    eng.reset(xs)
    eng.rollout(20)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize()
    e0.record()
    eng.rollout(100)
    e1.record()
    torch.cuda.synchronize()
    per = {}
    for _ in range(5):
        for name, ms in eng.profile_step():
            per[name] = per.get(name, 0.0) + ms / 5
    print(f"C3 FGNN_STEP_MODE={mode}: {e0.elapsed_time(e1) / 100 * 1e3:.1f} us/step | " + " ".join(f"{k}={v * 1e3:.1f}" for k, v in per.items()), flush=True)
    eng.close()
