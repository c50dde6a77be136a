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


def measure(n, b=1, k=3, hidden=32, radius=1.0, steps=200, warm=20, readout=0):
    xs = np.concatenate([make_workload(n, seed=11 + e) for e in range(min(b, 4))])
    if b > 4:
        xs = np.concatenate([xs] * ((b + 3) // 4))[:b * n]
    sd, _ = make_weights(hidden, k, 2)
    cap = int(max(24, 3.2 * np.pi * radius ** 2 * 1.6 + 16))
    eng = FlockEngine(n_agents=n, n_episodes=b, k=k, hidden=hidden, n_layers=2, comm_radius=radius, dt=0.01,
                      edge_capacity=cap, readout_mode=readout)
    eng.load_state_dict(sd)
    eng.reset(xs)
    eng.rollout(warm)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    eng.rollout(steps)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / steps
    st = eng.stats()
    assert not st["overflow"]
    d = st["n_edges"] / (n * b)
    eng.close()
    return n * b / (ms * 1e-3), ms, d


def main():
    peak, _ = measured_peaks()
    rows = []
    rows.append(("C1 N=100 K=3 H=32 (cfg/dagger.cfg)", measure(100, steps=2000, warm=50)))
    rows.append(("C2 N=10k K=3 H=64", measure(10_000, hidden=64, steps=2000, warm=50)))
    rows.append(("C2' N=10k K=3 H=64 FFMA readout", measure(10_000, hidden=64, steps=2000, warm=50, readout=1)))
    rows.append(("C3 256 x N=1k K=3 H=32", measure(1000, b=256, steps=500, warm=20)))
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
