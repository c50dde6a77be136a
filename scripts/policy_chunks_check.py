"""fgnn_policy to a HOST buffer: chunked readout overlapped with its D2H copies against the single-launch path.

    python scripts/policy_chunks_check.py [N] [K]

Variants: one launch (FGNN_POLICY_CHUNKS=1), equal chunks (FGNN_POLICY_PIPE=0, FGNN_POLICY_CHUNKS=4), and the pipelined
form (FGNN_POLICY_PIPE=w: last hop inside each chunk's readout, first chunk one wave of the readout grid, later chunks w
waves).  Every variant must leave identical actions, identical aggregated z and an identical state after env.step; the
e2e loop time (select_action -> pinned host -> env.step) of each is printed next to it."""
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
k = int(sys.argv[2]) if len(sys.argv) > 2 else 3
x0 = make_workload(n)
sd, _ = make_weights(32, k, 2)
variants = [("single launch", 1, 0), ("4 equal chunks", 4, 0), ("pipe 1", 4, 1), ("pipe 2", 4, 2), ("pipe 3", 4, 3),
            ("pipe 4", 4, 4)]
out = {}
for name, chunks, pipe in variants:
    os.environ["FGNN_POLICY_CHUNKS"] = str(chunks)
    os.environ["FGNN_POLICY_PIPE"] = str(pipe)
    eng = FlockEngine(n_agents=n, k=k, hidden=32, n_layers=2, comm_radius=1.0, dt=0.01, edge_capacity=32)
    eng.load_state_dict(sd)
    eng.reset(x0)
    eng.rollout(5)
    act = torch.empty((n, 2), dtype=torch.float32, pin_memory=True).numpy()
    acts = []
    for _ in range(3):
        act[:] = np.nan
        eng.policy(out=act)
        acts.append(act.copy())
        acts.append(eng.get_aggregated().copy())
        acts.append(eng.get_action().copy())
        eng.env_step(act)
    torch.cuda.synchronize()
    best = 1e9
    for _ in range(3):
        t0 = time.perf_counter()
        for _ in range(30):
            eng.policy(out=act)
            eng.env_step(act)
        torch.cuda.synchronize()
        best = min(best, (time.perf_counter() - t0) / 30 * 1e3)
    out[name] = (acts, eng.get_state(), best)
    eng.close()
ref = out["single launch"]
ok = True
for name, _, _ in variants[1:]:
    same = all(np.array_equal(a, b) for a, b in zip(out[name][0], ref[0])) and np.array_equal(out[name][1], ref[1])
    ok = ok and same
    print(f"N={n} K={k} {name}: identical={same}  e2e {out[name][2]:.3f} ms/step  (single launch: {ref[2]:.3f} ms/step)", flush=True)
sys.exit(0 if ok else 1)
