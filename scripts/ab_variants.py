"""A/B the kernel variants of the closed-loop step on the bench workload (one GPU):
    python scripts/ab_variants.py [N] [steps]
Variants are selected per engine by environment variables read in fgnn_create (FGNN_STEP_MODE,
FGNN_ADJ_MODE, FGNN_LAST_HOP_SEPARATE, FGNN_SCAN_TWO_PASS, FGNN_PDL).  Every variant must leave the SAME state bit for bit after the same number of steps
(the sums run in the same order); the script checks that, then prints graph-replay ms/step and per-kernel times."""
import itertools
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from bench import make_workload, make_weights      # noqa: E402
from multiagent_gnn_policies_b200.engine import FlockEngine        # noqa: E402


def run(n, steps, env, x0, sd, k=3, hidden=32, readout_mode=0):
    for key in ("FGNN_ADJ_MODE", "FGNN_LAST_HOP_SEPARATE", "FGNN_SCAN_TWO_PASS", "FGNN_PDL", "FGNN_STEP_MODE", "FGNN_PR_MINB"):
        os.environ.pop(key, None)
    os.environ.update(env)
    eng = FlockEngine(n_agents=n, k=k, hidden=hidden, n_layers=2, comm_radius=1.0, dt=0.01, readout_mode=readout_mode)
    eng.load_state_dict(sd)
    eng.reset(x0)
    eng.rollout(40)
    state40 = eng.get_state()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    best = 1e9
    for _ in range(3):
        torch.cuda.synchronize()
        e0.record()
        eng.rollout(steps)
        e1.record()
        torch.cuda.synchronize()
        best = min(best, e0.elapsed_time(e1) / steps)
    per = {}
    for _ in range(5):
        for name, ms in eng.profile_step():
            per[name] = per.get(name, 0.0) + ms / 5
    assert not eng.stats()["overflow"]
    eng.close()
    return best, per, state40


def main():
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 1_000_000
    steps = int(sys.argv[2]) if len(sys.argv) > 2 else 300
    x0 = make_workload(n)
    sd, _ = make_weights(32, 3, 2)
    ref_state = None
    print(f"N={n} steps={steps}")
    # round-1 step (separate adjacency / hop kernels) first: it defines the reference bits
    variants = [{"FGNN_STEP_MODE": "0"}]
    variants.append({"FGNN_STEP_MODE": "1"})
    for extra in os.environ.get("FGNN_AB_EXTRA", "").split(";"):
        if extra:
            variants.append(dict({"FGNN_STEP_MODE": "1"}, **dict(kv.split("=") for kv in extra.split(","))))
    for env in variants:
        ms, per, st = run(n, steps, env, x0, sd)
        if ref_state is None:
            ref_state = st
        same = bool(np.array_equal(st, ref_state))
        kern = " ".join(f"{k_}={v * 1e3:.1f}" for k_, v in per.items())
        tag = " ".join(f"{k_[5:].lower()}={v}" for k_, v in env.items())
        print(f"{tag}: {ms * 1e3:.1f} us/step  {n / ms / 1e6:.3f}e9 agent-steps/s  "
              f"bit-identical={same}  [{kern}] sum={sum(per.values()) * 1e3:.1f}", flush=True)


if __name__ == "__main__":
    main()
