"""Debugging aid (a checker, hence under tests/): large-N engine run printed next to the edge-list oracle, step by step."""
import sys, numpy as np, torch
sys.path.insert(0, '.')
from oracle import flock_env, learner, sparse
from multiagent_gnn_policies_b200.engine import FlockEngine
n, R, hidden = 30000, 2.0, 64
torch.manual_seed(11)
sd = {"conv_layers.0.weight": torch.randn(hidden, 6, 3, 1) * 0.2, "conv_layers.0.bias": torch.randn(hidden) * 0.1,
      "conv_layers.1.weight": torch.randn(hidden, hidden, 1, 1) * 0.1, "conv_layers.1.bias": torch.randn(hidden) * 0.1,
      "conv_layers.2.weight": torch.randn(2, hidden, 1, 1) * 0.1, "conv_layers.2.bias": torch.randn(2) * 0.1}
sd = {k: v.numpy() for k, v in sd.items()}
layers = learner.weights_from_state_dict(sd)
x = flock_env.synthetic_state(n, seed=n, density=1.6)
eng = FlockEngine(n_agents=n, k=3, hidden=hidden, n_layers=2, comm_radius=R, dt=0.01, edge_capacity=64)
eng.load_state_dict(sd)
eng.reset(x)
sstate = None
for t in range(4):
    sv, deg, i, j = sparse.compute_helpers_sparse(x, R)
    a_net = sparse.network_csr(n, deg, i, j)
    sstate = sparse.SparseDelayState(sv, a_net, prev_state=sstate, k=3)
    z_o = sstate.aggregate()
    act_o = sparse.readout(layers, z_o)
    act = eng.policy().cpu().numpy()
    z = eng.get_aggregated()
    f = eng.get_features()
    err = np.abs(act - act_o).max(axis=1)
    w = int(err.argmax())
    print(f"t={t} act rel_inf {err.max()/np.abs(act_o).max():.3e} worst agent {w} deg {deg[w]} act {act[w]} vs {act_o[w]}")
    print("   feature abs err max", np.abs(f - sv.astype(np.float32)).max(), "at worst agent", np.abs(f[w] - sv[w].astype(np.float32)).max(), "feat", sv[w])
    for k in range(3):
        print(f"   z[{k}] abs err max {np.abs(z[k]-z_o[k]).max():.3e}  worst-agent err {np.abs(z[k][w]-z_o[k][w]).max():.3e}  |z| max {np.abs(z_o[k]).max():.3e}")
    x = flock_env.integrate(x, act_o, 0.01)
    eng.env_step(act_o)
