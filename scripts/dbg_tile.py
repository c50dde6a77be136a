"""Debug aid: the warp-tile adjacency kernel k_pair_adjacency (FGNN_STEP_MODE=1) against the round-1 kernel k_adjacency_t
(FGNN_STEP_MODE=0), array by array (first written for the removed cell-tile kernel, hence the name)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from bench import make_workload, make_weights
from multiagent_gnn_policies_b200.engine import FlockEngine

def run(n, mode, steps, x0, sd):
    os.environ["FGNN_STEP_MODE"] = mode
    eng = FlockEngine(n_agents=n, k=3, hidden=32, n_layers=2, comm_radius=1.0, dt=0.01)
    eng.load_state_dict(sd)
    eng.reset(x0)
    out = []
    for t in range(steps):
        a = np.empty((n, 2), np.float32)
        d, f = eng.get_degrees(), eng.get_features()
        eng.policy(out=a)
        z = eng.get_aggregated()
        out.append((d, f, z, a.copy(), eng.get_state()))
        eng.env_step(a)
    eng.close()
    return out

for n in [int(v) for v in (sys.argv[1:] or ["100", "3000", "20000"])]:
    x0 = make_workload(n)
    sd, _ = make_weights(32, 3, 2)
    A, B = run(n, "0", 5, x0, sd), run(n, "1", 5, x0, sd)
    for t, (a, b) in enumerate(zip(A, B)):
        names = ["deg", "features", "z", "action", "state"]
        msg = []
        for nm, u, v in zip(names, a, b):
            bad = np.nonzero(np.any((u != v).reshape(len(u) if nm != "z" else 3, -1), axis=1))[0] if nm != "z" else np.nonzero(np.any((u != v).reshape(3, n, -1), axis=(0, 2)))[0]
            msg.append(f"{nm}:{bad.size}" + (f"{bad[:4].tolist()}" if bad.size else ""))
        print(f"n={n} t={t} mismatching agents -> " + "  ".join(msg), flush=True)
