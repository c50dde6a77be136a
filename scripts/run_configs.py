"""Measure the BASELINE.json configs C1..C5 on one GPU and print a markdown table.
    python scripts/run_configs.py            (GPU box; a few seconds per row)
Throughput is device-resident CUDA-graph rollout time (CUDA events), agent-steps/s = B*N*steps/s."""
import os
import sys
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from bench import make_workload, make_weights, measured_peaks      # noqa: E402
from multiagent_gnn_policies_b200.engine import FlockEngine        # noqa: E402


def measure(n, b=1, k=3, hidden=32, radius=1.0, steps=200, warm=20, readout=0, episodes=1):
    """`episodes` rollouts of `steps` timed steps each (reset + `warm` steps untimed in between): small flocks
    disperse within a few hundred free-running steps, so they are timed over episodes of the reference's length
    (TimeLimit 200) instead of one long run on an emptying graph."""
    xs = np.concatenate([make_workload(n, seed=11 + e) for e in range(min(b, 4))])
    if b > 4:
        xs = np.concatenate([xs] * ((b + 3) // 4))[:b * n]
    sd, _ = make_weights(hidden, k, 2)
    cap = int(max(24, 3.2 * np.pi * radius ** 2 * 1.6 + 16))
    eng = FlockEngine(n_agents=n, n_episodes=b, k=k, hidden=hidden, n_layers=2, comm_radius=radius, dt=0.01,
                      edge_capacity=cap, readout_mode=readout)
    eng.load_state_dict(sd)
    total_ms, d0, d1 = 0.0, 0.0, 0.0
    for ep in range(episodes):
        eng.reset(xs)
        eng.rollout(warm)
        d0 = eng.stats()["n_edges"] / (n * b)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        eng.rollout(steps)
        e1.record()
        torch.cuda.synchronize()
        total_ms += e0.elapsed_time(e1)
        st = eng.stats()
        assert not st["overflow"]
        d1 = st["n_edges"] / (n * b)
    ms = total_ms / (steps * episodes)
    eng.close()
    return n * b / (ms * 1e-3), ms, 0.5 * (d0 + d1)


def main():
    peak, _ = measured_peaks()
    rows = []
    rows.append(("C1 N=100 K=3 H=32 (cfg/dagger.cfg)", measure(100, steps=190, warm=10, episodes=10)))
    rows.append(("C2 N=10k K=3 H=64", measure(10_000, hidden=64, steps=190, warm=10, episodes=10)))
    rows.append(("C2' N=10k K=3 H=64 FFMA readout", measure(10_000, hidden=64, steps=190, warm=10, episodes=10, readout=1)))
    rows.append(("C3 256 x N=1k K=3 H=32", measure(1000, b=256, steps=190, warm=10, episodes=3)))
    for r in (0.8, 1.0, 1.5, 2.0, 3.0, 4.0):
        rows.append((f"C4 N=100k K=3 R={r}", measure(100_000, radius=r, steps=200, warm=20)))
    rows.append(("C5 N=1M K=3 H=32 (1 GPU)", measure(1_000_000, steps=100, warm=10)))
    rows.append(("N=1M K=3 H=64 (1 GPU)", measure(1_000_000, hidden=64, steps=100, warm=10)))
    rows.append(("N=4M K=3 H=32 (1 GPU)", measure(4_000_000, steps=50, warm=5)))
    print("| config | agent-steps/s | ms/step | mean degree | fraction of HBM roofline (268+12d B/agent-step) |")
    print("|---|---|---|---|---|")
    for name, (v, ms, d) in rows:
        frac = (268 + 12 * d) * v / 1e9 / peak
        print(f"| {name} | {v:.3e} | {ms:.4f} | {d:.2f} | {frac:.3f} |")


if __name__ == "__main__":
    main()
