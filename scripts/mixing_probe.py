"""How fast do velocities align and how far do agents wander relative to the flock (N=1M, trained policy)?"""
import os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from bench import make_workload, make_weights
from multiagent_gnn_policies_b200.engine import FlockEngine
N = 1_000_000
x0 = make_workload(N, seed=11)
sd, _ = make_weights(32, 3, 2)
eng = FlockEngine(n_agents=N, k=3, hidden=32, n_layers=2, comm_radius=1.0, dt=0.01, edge_capacity=32)
eng.load_state_dict(sd)
eng.reset(x0)
mean_v = x0[:, 2:4].mean(axis=0)
done = 0
for T in (0, 50, 100, 200, 400, 800, 1200):
    if T > done:
        r = eng.rollout(T - done, want_reward=True)
        done = T
        rew = r[-1, 0]
    else:
        rew = float('nan')
    x = eng.get_state()
    rel = x[:, 0:2] - x0[:, 0:2] - mean_v * (T * 0.01)
    print(f"t={T:5d} reward {rew:8.3f} vel std {x[:,2].std():.3f},{x[:,3].std():.3f}  rel displacement: rms {np.sqrt((rel**2).sum(1).mean()):.2f} max|dx| {np.abs(rel[:,0]).max():.2f}  p99.9|dx| {np.quantile(np.abs(rel[:,0]),0.999):.2f}  mean deg {eng.stats()['n_edges']/N:.2f}")
