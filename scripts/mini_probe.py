"""Small-flock timing: fused steps per second of the single-CTA path (fgnn_mini.cu) against the general kernels.
    python scripts/mini_probe.py"""
import os
import sys
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from bench import make_workload, make_weights      # noqa: E402
from multiagent_gnn_policies_b200.engine import FlockEngine        # noqa: E402

sd, _ = make_weights(32, 3, 2)
for n, b in ((100, 1), (50, 1), (128, 1), (32, 4)):
    for mini in ("0", "1"):
        os.environ["FGNN_MINI"] = mini
        eng = FlockEngine(n_agents=n, n_episodes=b, k=3, hidden=32, n_layers=2, comm_radius=1.0, dt=0.01)
        eng.load_state_dict(sd)
        x0 = np.concatenate([make_workload(n, seed=5 + e) for e in range(b)])
        eng.reset(x0)
        eng.rollout(50)
        best = 1e9
        for _ in range(5):
            eng.reset(x0)
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            eng.rollout(200)
            e1.record()
            torch.cuda.synchronize()
            best = min(best, e0.elapsed_time(e1) / 200)
        # host-driven single steps (fgnn_step with host action / reward out)
        a, r = np.empty((n * b, 2), np.float32), np.empty(b)
        eng.reset(x0)
        t0 = time.perf_counter()
        for _ in range(300):
            eng.step(a, r)
            eng.sync()
        step_us = (time.perf_counter() - t0) / 300 * 1e6
        eng.reset(x0)
        t0 = time.perf_counter()
        for _ in range(300):
            eng.policy(out=a)                            # select_action -> host (synchronises)
            eng.env_step(a)                              # env.step(host action) -> reward (synchronises)
        split_us = (time.perf_counter() - t0) / 300 * 1e6
        print(f"N={n} B={b} FGNN_MINI={mini}: policy + env_step {split_us:.1f} us | rollout {best * 1e3:.2f} us/step ({n * b / best / 1e3:.2f}e6 agent-steps/s) | "
              f"fgnn_step + sync {step_us:.1f} us", flush=True)
        eng.close()
