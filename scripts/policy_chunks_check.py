"""fgnn_policy to a HOST buffer: chunked readout overlapped with its D2H copies (FGNN_POLICY_CHUNKS) against the
single-launch path: identical actions, and the e2e loop time of both (select_action -> host -> env.step)."""
import os
import sys
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from bench import make_workload, make_weights      # noqa: E402
from multiagent_gnn_policies_b200.engine import FlockEngine        # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 1_000_000
x0 = make_workload(n)
sd, _ = make_weights(32, 3, 2)
out = {}
for chunks in (1, 4, 2):
    os.environ["FGNN_POLICY_CHUNKS"] = str(chunks)
    eng = FlockEngine(n_agents=n, k=3, hidden=32, n_layers=2, comm_radius=1.0, dt=0.01, edge_capacity=32)
    eng.load_state_dict(sd)
    eng.reset(x0)
    eng.rollout(5)
    act = torch.empty((n, 2), dtype=torch.float32, pin_memory=True).numpy()
    rew = np.empty(1, np.float64)
    acts = []
    for _ in range(3):
        eng.policy(out=act)
        acts.append(act.copy())
        eng.env_step(act)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(30):
        eng.policy(out=act)
        eng.env_step(act)
    torch.cuda.synchronize()
    ms = (time.perf_counter() - t0) / 30 * 1e3
    out[chunks] = (acts, eng.get_state(), ms)
    eng.close()
ref = out[1]
for chunks in (4, 2):
    same = all(np.array_equal(a, b) for a, b in zip(out[chunks][0], ref[0])) and np.array_equal(out[chunks][1], ref[1])
    print(f"chunks={chunks}: identical={same}  e2e {out[chunks][2]:.3f} ms/step  (single launch: {ref[2]:.3f} ms/step)")
