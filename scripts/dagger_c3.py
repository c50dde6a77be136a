"""BASELINE config C3 as a real DAGGER training rollout on one GPU: B = 256 parallel episodes x N = 1000 agents,
K = 3, expert labels + beta-mixed actions + device replay of aggregated features + native gradient steps.
    python scripts/dagger_c3.py [B] [N] [steps per episode] [updates per episode] [episodes]"""
import os
import sys
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from bench import make_workload      # noqa: E402
from multiagent_gnn_policies_b200.dagger import DeviceDagger        # noqa: E402


def main():
    a = [int(v) for v in sys.argv[1:]]
    B, N, steps, updates, episodes = (a + [256, 1000, 200, 200, 3][len(a):])[:5]
    xs = np.concatenate([make_workload(N, seed=11 + e) for e in range(min(B, 8))])
    x0 = np.concatenate([xs] * ((B + 7) // 8))[:B * N]
    dg = DeviceDagger(N, B, k=3, hidden=32, n_layers=2, lr=5e-5, buffer_steps=steps, batch_size=20, edge_capacity=32)
    for ep in range(episodes):
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        ret, _ = dg.run_episode(x0, steps=steps, updates=0)
        torch.cuda.synchronize()
        t1 = time.perf_counter()
        losses = [dg.gradient_step() for _ in range(updates)]
        dg._push_weights()
        torch.cuda.synchronize()
        t2 = time.perf_counter()
        print(f"episode {ep}: rollout {B}x{N}x{steps} in {(t1 - t0) * 1e3:.1f} ms = {B * N * steps / (t1 - t0):.3e} agent-steps/s "
              f"(expert + learner + store every step); {updates} gradient steps on {dg.batch_size * N} rows in "
              f"{(t2 - t1) * 1e3:.1f} ms = {(t2 - t1) / max(updates, 1) * 1e3:.3f} ms/step; loss {losses[0]:.4f} -> {losses[-1]:.4f}; "
              f"mean return {ret:.2f}", flush=True)
    dg.close()


if __name__ == "__main__":
    main()
