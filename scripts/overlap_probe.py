"""How much do independent flocks overlap on one GPU?  S engines of N/S agents each, one stream per engine, graph replays
interleaved, against one engine of N agents.  Upper bound for what a chunk-pipelined step could gain.
    python scripts/overlap_probe.py [N] [steps]"""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from bench import make_workload, make_weights      # noqa: E402
from multiagent_gnn_policies_b200.engine import FlockEngine        # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 1_000_000
steps = int(sys.argv[2]) if len(sys.argv) > 2 else 200
sd, _ = make_weights(32, 3, 2)
for S in (1, 2, 4, 8):
    streams = [torch.cuda.Stream() for _ in range(S)]
    engs = []
    for i in range(S):
        e = FlockEngine(n_agents=n // S, k=3, hidden=32, n_layers=2, comm_radius=1.0, dt=0.01, stream=streams[i].cuda_stream)
        e.load_state_dict(sd)
        e.reset(make_workload(n // S))
        e.rollout(10)
        engs.append(e)
    torch.cuda.synchronize()
    best = 1e9
    for _ in range(3):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        torch.cuda.synchronize()
        e0.record()
        for s_ in streams:
            s_.wait_event(e0)
        for _ in range(steps // 10):
            for e in engs:
                e.rollout(10)
        for s_ in streams:
            ev = torch.cuda.Event()
            ev.record(s_)
            torch.cuda.current_stream().wait_event(ev)
        e1.record()
        torch.cuda.synchronize()
        best = min(best, e0.elapsed_time(e1) / (steps // 10 * 10))
    print(f"S={S} engines x {n // S} agents: {best * 1e3:.1f} us per step of all  -> {n / best / 1e6:.3f}e9 agent-steps/s", flush=True)
    for e in engs:
        e.close()
