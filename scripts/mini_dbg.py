"""Debug aid: the single-CTA path at N = 100 over a range of edge-capacity hints (shared-memory sizing of its adjacency stage)."""
import sys
sys.path.insert(0, "/root/repo")
import numpy as np
from bench import make_workload, make_weights
from multiagent_gnn_policies_b200.engine import FlockEngine
sd, _ = make_weights(32, 3, 2)
for cap in (0, 48, 56, 64, 99):
    try:
        eng = FlockEngine(n_agents=100, k=3, hidden=32, n_layers=2, comm_radius=1.0, dt=0.01, edge_capacity=cap)
        eng.load_state_dict(sd); eng.reset(make_workload(100))
        a = np.empty((100, 2), np.float32)
        eng.policy(out=a); eng.env_step(a); eng.rollout(3)
        print("cap", cap, "ok", eng.stats()["step"])
        eng.close()
    except Exception as e:
        print("cap", cap, "FAIL", str(e)[:200])
