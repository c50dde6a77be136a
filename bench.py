#!/usr/bin/env python
"""Benchmark of the flocking-GNN rollout hot path (BASELINE.json metric: agent-steps/sec).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N ... bench.py --gpus N ...

One "step" = one closed-loop rollout step over every agent: radius adjacency + 6-d features ->
K-hop aggregation -> MLP readout -> double-integrator update.
Workload (config.workload): N agents PER GPU, K=3, uniform density 1.6 agents/unit^2 (mean degree ~5
at comm_radius 1), cell-major agent order, shipped-checkpoint-shaped actor (H=32, L=2), fp32 learner
arithmetic / fp64 environment arithmetic, synthetic data, random-init or shipped weights.

Prints ONE JSON line (rank 0).  `value` = device-resident throughput (CUDA-graph rollout);
`e2e` = the same metric through the reference-facing calls with HOST buffers every step
(select_action -> host, env.step(action from host)); `roofline` = dominant kernel vs measured HBM
peak; `cpu_baseline` = the reference's own learner code (oracle/_ref, unmodified) timed on this box's cores.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "agent-steps/sec"
UNIT = "agent-steps/s"
DENSITY = 1.6
FALLBACK_HBM_GBS = 6650.0      # /opt/skills/guides/B200_PROFILING.md fallback


def make_workload(n_agents, seed=11, density=DENSITY, v_max=3.0, min_dist=0.1, x_offset=0.0, bias_seed=None):
    """SURVEY.md 8(d) synthetic state: uniform square of side sqrt(N/density), cell-major order,
    no pair closer than min_dist (gym_flock's reset threshold)."""
    from scipy.spatial import cKDTree
    rng = np.random.default_rng(seed)
    side = np.sqrt(n_agents / density)
    x = np.empty((n_agents, 4))
    x[:, 0:2] = rng.uniform(0.0, side, size=(n_agents, 2))
    for _ in range(200):
        pairs = cKDTree(x[:, 0:2]).query_pairs(min_dist, output_type="ndarray")
        if pairs.size == 0:
            break
        bad = np.unique(pairs[:, 1])
        x[bad, 0:2] = rng.uniform(0.0, side, size=(bad.size, 2))
    # one flock = one common velocity bias: strips of a sharded flock pass the same bias_seed
    bias = (rng if bias_seed is None else np.random.default_rng(bias_seed)).uniform(-v_max, v_max, size=(2,))
    x[:, 2:4] = rng.uniform(-v_max, v_max, size=(n_agents, 2)) + bias
    order = np.lexsort((np.floor(x[:, 0]).astype(np.int64), np.floor(x[:, 1]).astype(np.int64)))
    x = x[order]
    x[:, 0] += x_offset
    return np.ascontiguousarray(x)


def make_weights(hidden, k, n_layers, seed=11):
    """Shipped checkpoint shape when available in the golden fixtures (H=32,K=3,L=2), else random init of the
    reference architecture (nn.Conv2d default init, torch.manual_seed(seed))."""
    if hidden == 32 and k == 3 and n_layers == 2:
        path = os.path.join(ROOT, "tests", "golden", "ckpt_n100_k3.npz")
        if os.path.exists(path):
            g = np.load(path)
            return {key[3:]: g[key] for key in g.files if key.startswith("sd.")}, "shipped checkpoint (via tests/golden)"
    import torch
    torch.manual_seed(seed)
    dims = [6] + [hidden] * n_layers + [2]
    sd = {}
    for i in range(len(dims) - 1):
        conv = torch.nn.Conv2d(dims[i], dims[i + 1], (k if i == 0 else 1, 1))
        sd[f"conv_layers.{i}.weight"] = conv.weight.detach().numpy()
        sd[f"conv_layers.{i}.bias"] = conv.bias.detach().numpy()
    return sd, "random init"


class ClockSampler:
    """nvidia-smi clock / throttle-reason sampling during the timed region."""
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index = index
        self.lines = []
        self.proc = None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                          "-lms", "20", "-i", str(self.index)], stdout=subprocess.PIPE,
                                         stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._read, daemon=True)
            self.thread.start()
        except OSError:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.lines.append(line.strip())

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for ln in self.lines:
            parts = [q.strip() for q in ln.split(",")]
            if len(parts) < 7:
                continue
            try:
                sm.append(float(parts[0]))
                mx.append(float(parts[1]))
            except ValueError:
                continue
            for nm, val in zip(names, parts[3:7]):
                if val.lower().startswith("active"):
                    reasons.add(nm)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "samples": len(sm), "reasons": sorted(reasons)}


NOT_STEP_SOURCES = ("fgnn_dense.cu",)      # dense Actor.forward for ind_agg > 0 (compat surface): no kernel of the step lives there


def kernel_source_sha():
    """sha1 over the CUDA sources of the step (every .cu / .cuh under csrc/ except NOT_STEP_SOURCES): profiles/r2_traffic.json
    records the build its ncu capture was taken from."""
    import hashlib
    h = hashlib.sha1()
    csrc = os.path.join(ROOT, "multiagent_gnn_policies_b200", "csrc")
    for name in sorted(os.listdir(csrc)):
        if name.endswith((".cu", ".cuh")) and name not in NOT_STEP_SOURCES:
            h.update(open(os.path.join(csrc, name), "rb").read())
    return h.hexdigest()[:16]


def ncu_traffic(kernel):
    """dram__bytes_read.sum + dram__bytes_write.sum per launch of `kernel` from the committed `ncu --set full` capture of
    this workload (profiles/r2_traffic.json, written by scripts/ncu_summary.py), or None when the capture is of another
    build of the kernels (the file records the source hash it was taken from)."""
    path = os.path.join(ROOT, "profiles", "r2_traffic.json")
    try:
        t = json.load(open(path))
        if t.get("_kernel_source_sha") != kernel_source_sha():
            return None
        return t.get(kernel)
    except Exception:
        return None


def small_config_c1(sd, local_rank, steps=2000):
    """BASELINE config C1 (cfg/dagger.cfg: N = 100, K = 3, H = 32) on this engine: the size the reference itself runs, for a
    like-for-like ratio against the N = 100 cell of `cpu_baseline` (agent-steps/s, device-resident and through host buffers)."""
    import torch
    from multiagent_gnn_policies_b200.engine import FlockEngine
    n = 100
    eng = FlockEngine(n_agents=n, k=3, hidden=32, n_layers=2, comm_radius=1.0, dt=0.01, device=local_rank)
    eng.load_state_dict(sd)
    eng.reset(make_workload(n, seed=11))
    eng.rollout(50)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize()
    e0.record()
    eng.rollout(steps)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / steps
    act = torch.empty((n, 2), dtype=torch.float32, pin_memory=True).numpy()
    rew = np.empty(1, np.float64)
    for _ in range(5):
        eng.policy(out=act)
        eng.lib.fgnn_env_step(eng._h, act.ctypes.data, rew.ctypes.data, eng.stream)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(200):
        eng.policy(out=act)
        eng.lib.fgnn_env_step(eng._h, act.ctypes.data, rew.ctypes.data, eng.stream)
        torch.cuda.current_stream().synchronize()
    e2e_ms = (time.perf_counter() - t0) / 200 * 1e3
    eng.close()
    return {"workload": "C1: FlockingRelative N=100 K=3 H=32 (cfg/dagger.cfg)", "value": n / (ms * 1e-3), "unit": UNIT,
            "us_per_step": ms * 1e3, "e2e_value": n / (e2e_ms * 1e-3), "e2e_us_per_step": e2e_ms * 1e3}


def measured_peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        try:
            return float(json.load(open(path))["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
        except Exception:
            pass
    return FALLBACK_HBM_GBS, "fallback (B200_PROFILING.md)"


# ---------------------------------------------------------------------------------------------
# reference arm / cpu baseline: the reference's OWN CPU implementation of the path --
# MultiAgentStateWithDelay (learner/state_with_delay.py:6-53) + DAGGER.select_action
# (learner/gnn_dagger.py:55-72) + Actor.forward (learner/actor.py:45-86), UNMODIFIED, from oracle/_ref/
# (staged by `python -m oracle.make_ref`; kind = "reference") -- driven by the spec env (oracle/flock_env.py:
# gym_flock is not installed anywhere here).  Without oracle/_ref the numpy port of the same dense algorithm
# (oracle/learner.py) is timed instead (kind = "port").  The dense algorithm holds 2 K N^2 floats per state, so
# it is timed on bounded samples of the workload (same density): N = 100 (the reference's own cfg/dagger.cfg)
# and N = 1000; `value` is the BEST agent-steps/s over samples and thread counts {1, all host cores}.
# ---------------------------------------------------------------------------------------------
REF_DIR = os.path.join(ROOT, "oracle", "_ref")


def _reference_learner(n_agents, hidden, k, n_layers, sd):
    """The unmodified reference DAGGER object (CPU) with the bench weights, or None if oracle/_ref is absent."""
    if not os.path.exists(os.path.join(REF_DIR, "learner", "gnn_dagger.py")):
        return None
    import configparser
    import importlib
    import torch
    if REF_DIR not in sys.path:
        sys.path.insert(0, REF_DIR)
    gd = importlib.import_module("learner.gnn_dagger")
    swd = importlib.import_module("learner.state_with_delay")
    cp = configparser.ConfigParser()
    cp.read(os.path.join(REF_DIR, "cfg", "dagger.cfg"))
    args = cp[cp.default_section]
    args["n_agents"], args["hidden_size"], args["k"], args["n_layers"] = str(n_agents), str(hidden), str(k), str(n_layers)
    device = torch.device("cpu")
    learner = gd.DAGGER(device, args)
    learner.actor.load_state_dict({key: torch.as_tensor(np.asarray(v)) for key, v in sd.items()})
    return learner, swd.MultiAgentStateWithDelay, args, device


def _time_reference(n_sample, steps, warmup, hidden, k, n_layers, threads, seed=11):
    """(seconds per step, kind) of the host loop of learner/gnn_dagger.py:196-201 at N = n_sample with `threads` threads."""
    import torch
    from oracle import flock_env, learner as port
    try:
        from threadpoolctl import threadpool_limits
    except Exception:
        threadpool_limits = None
    sd, _ = make_weights(hidden, k, n_layers)
    torch.set_num_threads(int(threads))
    ref = _reference_learner(n_sample, hidden, k, n_layers, sd)
    x = make_workload(n_sample, seed=seed)
    R2, dt_env = 1.0, 0.01
    if ref is not None:
        dagger, State, args, device = ref
        kind = "reference"

        def one_step(x, state):
            sv, sn, _, _ = flock_env.compute_helpers(x, R2)                         # env side: spec env (gym_flock absent)
            state = State(device, args, (sv, sn), prev_state=state)                # state_with_delay.py:6-53
            a = dagger.select_action(state).cpu().numpy()                          # gnn_dagger.py:55-72, :161
            return flock_env.integrate(x, a, dt_env), state
    else:
        layers = port.weights_from_state_dict(sd)
        kind = "port"

        def one_step(x, state):
            sv, sn, _, _ = flock_env.compute_helpers(x, R2)
            state = port.DelayState((sv, sn), prev_state=state, k=k, with_curr_gso=True)
            return flock_env.integrate(x, port.select_action(layers, state), dt_env), state

    def loop():
        xx, state = x, None
        for _ in range(warmup):
            xx, state = one_step(xx, state)
        t0 = time.perf_counter()
        for _ in range(steps):
            xx, state = one_step(xx, state)
        return (time.perf_counter() - t0) / steps

    if threadpool_limits is not None:
        with threadpool_limits(limits=int(threads)):
            return loop(), kind
    return loop(), kind


def run_cpu_reference(samples, steps, warmup, hidden, k, n_layers, seed=11):
    """Times the reference path on each N of `samples` with 1 thread and with every host core; returns
    (cpu_baseline dict, seconds per step of the best cell, N of the best cell)."""
    ncpu = os.cpu_count() or 1
    cells, best = [], None
    for n in samples:
        for th in sorted({1, ncpu}):
            sec, kind = _time_reference(n, steps, warmup, hidden, k, n_layers, th, seed)
            cell = {"n_agents": n, "threads": th, "ms_per_step": sec * 1e3, "agent_steps_per_s": n / sec}
            cells.append(cell)
            if best is None or cell["agent_steps_per_s"] > best["agent_steps_per_s"]:
                best = cell
    what = ("UNMODIFIED reference learner from oracle/_ref (MultiAgentStateWithDelay + DAGGER.select_action, torch CPU)"
            if kind == "reference" else "numpy port of the reference's dense algorithm (oracle/learner.py; oracle/_ref absent)")
    sample = (f"{what} + spec env (oracle/flock_env.py, float64 all-pairs); best cell: N={best['n_agents']} agents of the "
              f"same density, {best['threads']} thread(s), {steps} steps after {warmup} warm-up, "
              f"{best['ms_per_step']:.2f} ms/step; cells tried (N, threads, agent-steps/s): "
              + ", ".join(f"({c['n_agents']}, {c['threads']}, {c['agent_steps_per_s']:.3g})" for c in cells)
              + f"; host has {ncpu} cpus")
    return ({"value": best["agent_steps_per_s"], "unit": UNIT, "cores": int(best["threads"]), "kind": kind,
             "sample": sample, "cells": cells}, best["ms_per_step"] * 1e-3, best["n_agents"])


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=1000)
    ap.add_argument("--warmup", type=int, default=50)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--n-agents", type=int, default=1_000_000, help="agents per GPU")
    ap.add_argument("--hidden", type=int, default=32)
    ap.add_argument("--k", type=int, default=3)
    ap.add_argument("--n-layers", type=int, default=2)
    ap.add_argument("--radius", type=float, default=1.0)
    ap.add_argument("--cpu-sample", type=int, default=0, help="N of the bounded CPU-baseline sample (0: N=100 and N=1000)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--e2e-steps", type=int, default=0, help="steps of the host-buffer e2e loop (default: min(steps, 50))")
    args = ap.parse_args()

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    workload = (f"FlockingRelative closed-loop rollout, N={args.n_agents} agents per GPU, K={args.k}, H={args.hidden}, "
                f"L={args.n_layers}, R={args.radius}, density {DENSITY}/unit^2, dt=0.01, cell-major order")
    config = {"workload": workload, "n_agents_per_gpu": args.n_agents, "k": args.k, "hidden": args.hidden,
              "n_layers": args.n_layers, "comm_radius": args.radius,
              "l2_policy": "working set per step (~0.5 GB at N=1M) exceeds the 126 MB L2; no explicit flush"}

    if args.impl == "reference":
        if rank != 0:
            return
        # exactly --steps timed steps after --warmup untimed ones, per cell (sample size x thread count); a run at the
        # driver's --steps 20 --warmup 5 takes well under a minute of host time
        steps, warm = max(1, args.steps), max(0, args.warmup)
        samples = [args.cpu_sample] if args.cpu_sample > 0 else [100, 1000]
        cb, sec, n_best = run_cpu_reference(samples, steps, warm, args.hidden, args.k, args.n_layers)
        config = dict(config, workload=f"reference sample N={n_best} agents (dense reference path, cells tried N={samples}; the "
                                       f"dense (K,N,N) operators of N={args.n_agents} do not fit any host) of: " + workload)
        line = {"impl": "reference", "metric": METRIC, "value": cb["value"], "unit": UNIT, "n_gpus": args.gpus,
                "steps": steps, "warmup": warm, "ms_per_step": sec * 1e3, "higher_is_better": True, "scaling": "weak",
                "vs_baseline": None, "dtype": "f32 learner / f64 env", "data": "synthetic", "config": config,
                "cpu_baseline": cb, "gpu_launches": 0,
                "e2e": {"value": cb["value"], "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
        print(json.dumps(line))
        return

    import torch
    import torch.distributed as dist
    from multiagent_gnn_policies_b200.engine import FlockEngine

    torch.cuda.set_device(local_rank)
    if world > 1:
        # before any pinned allocation: the rank and its host buffers on the GPU's own NUMA node (FGNN_BIND_CPUS=0: leave it)
        if os.environ.get("FGNN_BIND_CPUS", "1") != "0":
            from multiagent_gnn_policies_b200 import parallel
            cpus = parallel.bind_to_local_cpus(local_rank)
            config["cpu_binding"] = f"rank 0 bound to {len(cpus)} GPU-local cpus" if cpus else "unchanged"
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))

    N = args.n_agents
    side = np.sqrt(N / DENSITY)
    sd, weights_note = make_weights(args.hidden, args.k, args.n_layers)
    cap = int(max(24, 3.2 * np.pi * args.radius ** 2 * DENSITY + 16))

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def max_over_ranks(ms):
        if world > 1:
            tms = torch.tensor([ms], device="cuda")
            dist.all_reduce(tms, op=dist.ReduceOp.MAX)
            return float(tms.item())
        return ms

    if world > 1:
        run_sharded(args, rank, world, local_rank, N, side, sd, weights_note, cap, config, barrier, max_over_ranks)
        dist.destroy_process_group()
        return

    x0 = make_workload(N, seed=11)
    eng = FlockEngine(n_agents=N, k=args.k, hidden=args.hidden, n_layers=args.n_layers, comm_radius=args.radius,
                      dt=0.01, device=local_rank, edge_capacity=cap)
    eng.load_state_dict(sd)
    eng.reset(x0)
    st0 = eng.stats()
    deg_start = st0["n_edges"] / N

    # ---- device-resident throughput: CUDA-graph rollout ------------------------------------
    eng.rollout(args.warmup)
    barrier()
    sampler = ClockSampler(local_rank)
    sampler.start()
    l0 = eng.launch_count()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    e0.record()
    eng.rollout(args.steps)
    e1.record()
    barrier()
    ms = e0.elapsed_time(e1)
    launches = eng.launch_count() - l0
    clocks = sampler.stop()
    st1 = eng.stats()
    deg_end = st1["n_edges"] / N
    if st1["overflow"]:
        raise RuntimeError("edge capacity overflow during the timed region: results void")
    value = world * N * args.steps / (ms * 1e-3)

    # ---- per-kernel device times (CUDA events between launches), for the roofline ----------
    prof = {}
    nprof = 10
    for _ in range(nprof):
        for name, kms in eng.profile_step():
            prof[name] = prof.get(name, 0.0) + kms / nprof
    deg_prof = eng.stats()["n_edges"] / N

    # ---- e2e: reference-facing calls with HOST buffers every step ---------------------------
    e2e_steps = args.e2e_steps or min(args.steps, 50)
    act_host = torch.empty((N, 2), dtype=torch.float32, pin_memory=True)
    rew_host = torch.empty((1,), dtype=torch.float64, pin_memory=True)
    act_np, rew_np = act_host.numpy(), rew_host.numpy()
    lib, h, stream = eng.lib, eng._h, eng.stream
    for _ in range(3):
        eng.policy(out=act_np)
        lib.fgnn_env_step(h, act_np.ctypes.data, rew_np.ctypes.data, stream)
    barrier()
    t0 = torch.cuda.Event(enable_timing=True)
    t1 = torch.cuda.Event(enable_timing=True)
    t0.record()
    for _ in range(e2e_steps):
        eng.policy(out=act_np)                       # select_action(state) -> host action (D2H, sync)
        rc = lib.fgnn_env_step(h, act_np.ctypes.data, rew_np.ctypes.data, stream)    # env.step(host action) (H2D) -> reward (D2H)
        assert rc == 0
        torch.cuda.current_stream().synchronize()    # the caller reads the reward
    t1.record()
    barrier()
    e2e_ms = t0.elapsed_time(t1)
    e2e_value = world * N * e2e_steps / (e2e_ms * 1e-3)

    # ---- roofline of the dominant kernel ---------------------------------------------------
    peak, peak_src = measured_peaks()
    # SURVEY.md 8(d) algorithmic bytes per agent per launch (K=3; fp32 values, int32 CSR, 16 B state)
    d = deg_prof
    alg_bytes = {"adjacency": 44 + 4 * d, "pair_adjacency": 44 + 4 * d, "hop0": 104 + 4 * d, "final": 120 + 4 * d}
    if "hop_last" in prof:
        # the last hop runs as its own launch: SURVEY's K3 row splits into the gather (CSR 4d+4, deg 4, source rows 24,
        # z_2 written 24) and the streaming readout + integrator (x_t, z_1, z_2 24 each, state 16 in / 16 out, action 8)
        alg_bytes["hop_last"] = 56 + 4 * d
        alg_bytes["final"] = 112
    dom = max(prof, key=prof.get)
    step_ms_prof = sum(prof.values())
    roof = {"bound": "hbm", "kernel": dom, "unit": "GB/s", "peak": peak, "peak_source": peak_src,
            "traffic": ncu_traffic(dom) if (N == 1_000_000 and args.k == 3 and args.hidden == 32) else None,
            "kernel_ms": prof[dom], "kernel_share_of_step": prof[dom] / step_ms_prof,
            "per_kernel_ms": {k_: round(v, 5) for k_, v in prof.items()}}
    if dom in alg_bytes and args.k == 3:
        ach = alg_bytes[dom] * N / (prof[dom] * 1e-3) / 1e9
        roof.update({"achieved": ach, "frac": ach / peak, "algorithmic_bytes_per_agent": alg_bytes[dom]})
    else:
        roof.update({"achieved": None, "frac": None})
    step_bytes = 268 + 12 * d
    roof["whole_step"] = {"algorithmic_bytes_per_agent_step": step_bytes,
                          "achieved": step_bytes * value / world / 1e9, "frac": step_bytes * value / world / 1e9 / peak}
    roof["per_kernel_frac"] = {k_: round(alg_bytes[k_] * N / (v * 1e-3) / 1e9 / peak, 4) for k_, v in prof.items()
                               if k_ in alg_bytes and args.k == 3}
    if dom in ("adjacency", "pair_adjacency"):
        roof["note"] = ("the dominant kernel is the float64 pair test + feature kernel: 64 algorithmic bytes per agent but "
                        "~14 candidate pairs per agent in float64 -- issue/fp64-pipe bound, not HBM bound (profiles/)")

    cb, equal_n = None, None
    if not args.no_cpu_baseline:
        cb, _, _ = run_cpu_reference([args.cpu_sample] if args.cpu_sample > 0 else [100, 1000], 5, 2, args.hidden, args.k,
                                     args.n_layers)
        if args.hidden == 32 and args.k == 3 and args.n_layers == 2:
            # like for like: BASELINE config C1 (the reference's own N = 100) on this engine against the N = 100 cell
            equal_n = small_config_c1(sd, local_rank)
            ref100 = max([c["agent_steps_per_s"] for c in cb["cells"] if c["n_agents"] == 100] or [0.0])
            if ref100 > 0:
                equal_n.update(reference_value=ref100, ratio_device=equal_n["value"] / ref100,
                               ratio_e2e=equal_n["e2e_value"] / ref100)

    line = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": ms / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f32 learner / f64 env", "data": f"synthetic ({weights_note})", "config": config,
            "mean_degree": {"start": deg_start, "end": deg_end},
            "clocks": clocks, "gpu_launches": int(launches),
            "e2e": {"value": e2e_value, "unit": UNIT, "steps": e2e_steps, "ms_per_step": e2e_ms / e2e_steps,
                    "h2d_bytes_per_step": int(N * 2 * 4), "d2h_bytes_per_step": int(N * 2 * 4 + 8)},
            "roofline": roof, "cpu_baseline": cb, "equal_n": equal_n}
    print(json.dumps(line))


def sharded_rollout(args, rank, world, local_rank, n_per_rank, sd, edge_cap, barrier, max_over_ranks, steps, warmup,
                    want_e2e=False, sample_clocks=False):
    """ONE flock of world * n_per_rank agents, sharded by index into x-strips (rank r owns the r-th strip), closed-loop
    rollout with the per-step halo exchange.  Returns a dict (identical on every rank up to the rank-0-only fields)."""
    import torch
    from multiagent_gnn_policies_b200 import parallel
    N = n_per_rank
    side = np.sqrt(N / DENSITY)
    n_total = world * N
    ranges = parallel.shard_ranges(n_total, world)
    lo, cnt = ranges[rank]
    # every rank generates its own strip and the two neighbouring strips (all it needs at reset)
    x_global = np.zeros((n_total, 4))
    x_global[:, 0] = parallel.FAR
    for q in (rank - 1, rank, rank + 1):
        if 0 <= q < world:
            x_global[ranges[q][0]:ranges[q][0] + ranges[q][1]] = make_workload(N, seed=11 + q, x_offset=q * side,
                                                                               bias_seed=11)
    depth = parallel.halo_depth(args.k, args.radius)
    # both boundaries of a strip; ownership hand-over keeps the layer thin for any run length
    halo_cap = int(1.5 * (depth + 2 * args.radius) * side * DENSITY * 2) + 1024
    cell = args.radius
    gx = int(np.ceil((side + 2 * depth + 4) / cell)) + 2
    gy = int(np.ceil(side / cell)) + 4
    be = parallel.CudaShardBackend(n_total, lo, cnt, ghost_capacity=2 * halo_cap, device=local_rank, k=args.k,
                                   hidden=args.hidden, n_layers=args.n_layers, comm_radius=args.radius, dt=0.01,
                                   edge_capacity=edge_cap, grid_dim=gx, grid_dim_y=gy)
    be.engine.load_state_dict(sd)
    flock = parallel.ShardedFlock(be, rank, world, args.k, args.radius, halo_cap,
                                  parallel.nccl_all_gather(world, halo_cap, be.device))
    bounds = np.array([-parallel.INF] + [q * side for q in range(1, world)] + [parallel.INF])
    frame_vx = float(np.random.default_rng(11).uniform(-3.0, 3.0, size=(2,))[0])       # the flock's common velocity bias
    flock.reset(x_global, ranges, bounds=bounds, frame_velocity=frame_vx)
    halo = os.environ.get("FGNN_HALO", "p2p")
    if halo == "p2p":
        # default: halo records stored straight into the peers' inboxes over NVLink (CUDA-IPC), one CUDA graph per step
        flock.enable_p2p(parallel.torch_all_gather_object(world))
        transport = "p2p stores into the peers' inboxes over NVLink (CUDA-IPC), one CUDA graph per step, no collective"
    elif halo == "native":
        be.init_comm(rank, world)               # ncclAllGather inside ONE step graph on the engine's own communicator
        transport = "ncclAllGather inside the step graph"
    else:
        transport = "NCCL all-gather via torch.distributed between two graph halves"
    for _ in range(warmup):
        flock.step()
    barrier()
    sampler = ClockSampler(local_rank)
    if rank == 0 and sample_clocks:
        sampler.start()
    l0 = be.engine.launch_count()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    e0.record()
    for _ in range(steps):
        flock.step()
    e1.record()
    barrier()
    ms = max_over_ranks(e0.elapsed_time(e1))
    launches = be.engine.launch_count() - l0
    clocks = sampler.stop() if (rank == 0 and sample_clocks) else None
    st = be.engine.stats()
    if st["overflow"]:
        raise RuntimeError(f"capacity overflow / halo time-out ({st['overflow']}) during the timed region: results void")
    if os.environ.get("FGNN_BENCH_PROFILE") == "1" and halo == "p2p":
        per = {}
        for _ in range(10):                      # the p2p step kernel by kernel (CUDA events, un-graphed), every rank in step
            for name, kms in be.engine.profile_step():
                per[name] = per.get(name, 0.0) + kms / 10
        print(f"[profile] rank {rank} N/gpu={N}: graph step {ms / steps * 1e3:.1f} us | "
              + " ".join(f"{k}={v * 1e3:.1f}" for k, v in per.items()), file=sys.stderr, flush=True)
    out = {"n_agents_total": n_total, "n_agents_per_gpu": N, "value": n_total * steps / (ms * 1e-3), "ms_per_step": ms / steps,
           "launches": int(launches), "clocks": clocks, "ghosts": int(st.get("n_ghosts", 0)), "owned": len(be.owned()),
           "mean_degree": st["n_edges"] / max(1, cnt + int(st.get("n_ghosts", 0))), "transport": transport,
           "halo_cap": flock.cap, "depth": flock.depth, "rows_io": be.engine.rows_io}
    if want_e2e:
        # e2e: the same step through host buffers (select_action -> pinned host -> env.step), halo over the same transport
        e2e_steps = args.e2e_steps or min(steps, 50)
        act_host = torch.empty((be.engine.rows_io, 2), dtype=torch.float32, pin_memory=True).numpy()
        for _ in range(3):
            flock.step_host(act_host)
        barrier()
        t0, t1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        t0.record()
        for _ in range(e2e_steps):
            flock.step_host(act_host)
        t1.record()
        barrier()
        e2e_ms = max_over_ranks(t0.elapsed_time(t1))
        out.update(e2e_steps=e2e_steps, e2e_ms_per_step=e2e_ms / e2e_steps, e2e_value=n_total * e2e_steps / (e2e_ms * 1e-3))
    be.engine.close()
    barrier()
    return out


def run_sharded(args, rank, world, local_rank, N, side, sd, weights_note, edge_cap, config, barrier, max_over_ranks):
    """N > 1.  Headline (`value`, scaling "weak"): ONE flock of world * N agents, N per GPU.  Next to it the fixed-size runs
    north_star names (`strong`): N agents IN TOTAL sharded over the GPUs (BASELINE config C5 at N = 1M) and config C4
    (100 000 agents in total, up to 4 GPUs)."""
    weak = sharded_rollout(args, rank, world, local_rank, N, sd, edge_cap, barrier, max_over_ranks, args.steps, args.warmup,
                           want_e2e=True, sample_clocks=True)
    strong = []
    for n_total, label in ((N, f"C5: N={N} agents in total"), (100_000, "C4: N=100000 agents in total")):
        if n_total % world or (n_total == 100_000 and world > 4) or n_total // world < 10_000:
            continue
        r = sharded_rollout(args, rank, world, local_rank, n_total // world, sd, edge_cap, barrier, max_over_ranks,
                            max(args.steps, 200), max(args.warmup, 20))
        strong.append({"workload": label, "n_agents_total": n_total, "n_agents_per_gpu": n_total // world, "value": r["value"],
                       "unit": UNIT, "ms_per_step": r["ms_per_step"], "halo_records_per_step": r["ghosts"]})
    if rank != 0:
        return
    peak, peak_src = measured_peaks()
    d = weak["mean_degree"]
    step_bytes = 268 + 12 * d
    value = weak["value"]
    config = dict(config, parallelism=f"index-sharded x{world} (x-strips), halo: {weak['transport']}; {weak['halo_cap']}-record inboxes, "
                                      f"halo depth {weak['depth']:.2f}, ownership hand-over", n_agents_total=weak["n_agents_total"])
    line = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": weak["ms_per_step"], "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f32 learner / f64 env", "data": f"synthetic ({weights_note})", "config": config,
            "halo_records_per_step": weak["ghosts"], "owned_by_rank0_after_run": weak["owned"], "clocks": weak["clocks"],
            "gpu_launches": weak["launches"],
            "e2e": {"value": weak["e2e_value"], "unit": UNIT, "steps": weak["e2e_steps"],
                    "ms_per_step": weak["e2e_ms_per_step"], "h2d_bytes_per_step": int(weak["rows_io"] * 2 * 4) * world,
                    "d2h_bytes_per_step": int(weak["rows_io"] * 2 * 4) * world},
            "strong": strong,
            "roofline": {"bound": "hbm", "kernel": None, "unit": "GB/s", "peak": peak, "peak_source": peak_src,
                         "achieved": step_bytes * value / world / 1e9, "frac": step_bytes * value / world / 1e9 / peak,
                         "traffic": None, "note": "whole step per GPU (268 + 12 d) B per agent-step; per-kernel breakdown "
                                                  "is reported by the 1-GPU run"},
            "cpu_baseline": None}
    print(json.dumps(line))


if __name__ == "__main__":
    main()
