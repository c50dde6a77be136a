"""How much do the gather kernels depend on the caller's index order?  Same flock, three index orders:
row-major cell order (the bench workload), 2-D tiled cell order (8x8-cell tiles), random permutation."""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from bench import make_workload, make_weights      # noqa: E402
from multiagent_gnn_policies_b200.engine import FlockEngine        # noqa: E402


def measure(x0, sd, steps=200):
    n = x0.shape[0]
    eng = FlockEngine(n_agents=n, k=3, hidden=32, n_layers=2, comm_radius=1.0, dt=0.01)
    eng.load_state_dict(sd)
    eng.reset(x0)
    eng.rollout(30)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize()
    e0.record()
    eng.rollout(steps)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / steps
    per = {}
    for _ in range(5):
        for name, v in eng.profile_step():
            per[name] = per.get(name, 0.0) + v / 5
    eng.close()
    return ms, per


def main():
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 1_000_000
    x = make_workload(n)
    sd, _ = make_weights(32, 3, 2)
    cx, cy = np.floor(x[:, 0]).astype(np.int64), np.floor(x[:, 1]).astype(np.int64)
    orders = {"row-major cells": np.arange(n)}
    for t in (4, 8, 16):
        orders[f"{t}x{t}-cell tiles"] = np.lexsort((cx, cy, cx // t, cy // t))
    orders["random permutation"] = np.random.default_rng(0).permutation(n)
    for name, o in orders.items():
        ms, per = measure(np.ascontiguousarray(x[o]), sd)
        kern = " ".join(f"{k_}={v * 1e3:.1f}" for k_, v in per.items())
        print(f"{name:22s} {ms * 1e3:7.1f} us/step {n / ms / 1e6:.3f}e9 agent-steps/s [{kern}]", flush=True)


if __name__ == "__main__":
    main()
