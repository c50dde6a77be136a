import os, sys
import numpy as np, torch
ROOT = "/root/repo"
sys.path.insert(0, ROOT)
from bench import make_workload, make_weights, DENSITY
from multiagent_gnn_policies_b200.engine import FlockEngine
n = 1_000_000
x0 = make_workload(n)
sd, _ = make_weights(32, 3, 2)
side = np.sqrt(n / DENSITY)
def prof(eng, reps=5):
    per = {}
    for _ in range(reps):
        for name, ms in eng.profile_step():
            per[name] = per.get(name, 0.0) + ms / reps
    return " ".join(f"{k}={v*1e3:.1f}" for k, v in per.items())
for kw in ({}, {"grid_dim": 803, "grid_dim_y": 794}, {"grid_dim": 803, "grid_dim_y": 794, "edge_capacity": 32}, {"grid_dim": 1200, "grid_dim_y": 1200}):
    eng = FlockEngine(n_agents=n, k=3, hidden=32, n_layers=2, comm_radius=1.0, dt=0.01, **kw)
    eng.load_state_dict(sd); eng.reset(x0); eng.rollout(30)
    print(kw, eng.stats()["grid_dim"], prof(eng), flush=True)
    eng.close()
