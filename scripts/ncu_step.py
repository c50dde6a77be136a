"""Tiny driver for ncu: a few closed-loop steps (no CUDA graph: each kernel is its own launch) on the bench workload.
    ncu --set full ... python scripts/ncu_step.py [N] [steps]
Kernel variants follow the FGNN_* environment variables."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from bench import make_workload, make_weights      # noqa: E402
from multiagent_gnn_policies_b200.engine import FlockEngine        # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 1_000_000
steps = int(sys.argv[2]) if len(sys.argv) > 2 else 6
sd, _ = make_weights(32, 3, 2)
eng = FlockEngine(n_agents=n, k=3, hidden=32, n_layers=2, comm_radius=1.0, dt=0.01)
eng.load_state_dict(sd)
eng.reset(make_workload(n))
for _ in range(steps):
    eng.step(None, None)
eng.sync()
print("steps done", eng.stats())
