import os, sys
sys.path.insert(0, "/root/repo")
os.environ["FGNN_MINI_CLOCK"] = "1"
import numpy as np
from bench import make_workload, make_weights
from multiagent_gnn_policies_b200.engine import FlockEngine
sd, _ = make_weights(32, 3, 2)
eng = FlockEngine(n_agents=100, k=3, hidden=32, n_layers=2, comm_radius=1.0, dt=0.01)
eng.load_state_dict(sd); eng.reset(make_workload(100)); eng.rollout(1000); print(eng.stats())
