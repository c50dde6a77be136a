"""Parity of the CUDA path (through the C ABI) against the golden vectors made by the reference's own
learner code, and against the CPU oracle.  Run on the B200 box: pytest -m gpu."""
import os

import numpy as np
import pytest

from conftest import ROOT, rel_inf, load_golden, golden_names
from oracle import flock_env, learner, sparse

pytestmark = pytest.mark.gpu

TOL_ACTION = 1e-5       # north_star: actions within 1e-5 relative (inf-norm) fp32


def log_parity(tag, act, act_oracle, truth):
    """The tests that compare at large N accept max(1e-5, c x the fp32 oracle's own error against float64): print (and, on
    the GPU box, append to gpurun_out/parity_ratios.log) the observed numbers so it is visible how often the relaxed branch
    decides."""
    e_cuda, e_oracle, e_pair = rel_inf(act, truth), rel_inf(act_oracle, truth), rel_inf(act, act_oracle)
    line = (f"[parity] {tag}: cuda-vs-f64 {e_cuda:.3e}  oracle(fp32)-vs-f64 {e_oracle:.3e}  cuda-vs-oracle {e_pair:.3e}  "
            f"strict(1e-5)={'pass' if max(e_cuda, e_pair) <= TOL_ACTION else 'RELAXED'}")
    print(line)
    out = os.path.join(ROOT, "gpurun_out")
    if os.path.isdir(out):
        with open(os.path.join(out, "parity_ratios.log"), "a") as f:
            f.write(line + "\n")


TOL_FEATURE = 2e-7      # fp32 rounding of a float64-accurate sum


def make_engine(g, **kw):
    from multiagent_gnn_policies_b200.engine import FlockEngine
    eng = FlockEngine(n_agents=g["n_agents"], k=g["k"], hidden=g["hidden"], n_layers=g["n_layers"],
                      comm_radius=g["comm_radius"], dt=g["dt"], **kw)
    eng.load_state_dict(g["state_dict"])
    return eng


FFMA, TENSOR = 1, 2      # fgnn_config.readout_mode
READOUTS = [FFMA, TENSOR]


@pytest.mark.parametrize("name", golden_names())
@pytest.mark.parametrize("mode", ["env_step", "teacher_forced"])
@pytest.mark.parametrize("readout", READOUTS)
def test_golden_trajectory(name, mode, readout):
    g = load_golden(name)
    if readout == TENSOR and g["hidden"] > 64:
        pytest.skip("tensor-core readout covers hidden <= 64")
    if readout == 3 and not 16 < g["hidden"] <= 64:
        pytest.skip("two-warp readout covers hidden in 17..64")
    eng = make_engine(g, readout_mode=readout)
    eng.reset(g["x"][0])
    for t in range(g["steps"]):
        if t > 0:
            if mode == "env_step":
                r = eng.env_step(g["action"][t - 1])
                assert r[0] == pytest.approx(float(g["reward"][t - 1]), rel=1e-9, abs=1e-12)
                np.testing.assert_array_equal(eng.get_state(), g["x"][t])     # integrator is bit-exact
            else:
                eng.set_state(g["x"][t])
                eng.build_graph(advance=True)
        assert np.array_equal(eng.get_degrees(), g["deg"][t])                 # adjacency: exact
        feats = eng.get_features()
        assert rel_inf(feats, g["values"][t].astype(np.float32)) <= TOL_FEATURE
        act = eng.policy().cpu().numpy()
        assert rel_inf(act, g["action"][t]) <= TOL_ACTION
        z = eng.get_aggregated()                                              # (K, N, 6)
        assert rel_inf(z.transpose(0, 2, 1), g["z"][t]) <= 2e-6
    assert not eng.stats()["overflow"]
    eng.close()


@pytest.mark.parametrize("name", ["ckpt_n100_k3", "rand_n100_k4_h64_l2", "rand_n50_k2_h16_l3", "rand_n12_k1_h4_l1"])
def test_network_export_equals_oracle(name):
    g = load_golden(name)
    eng = make_engine(g)
    eng.reset(g["x"][0])
    for t in range(min(3, g["steps"])):
        if t > 0:
            eng.env_step(g["action"][t - 1])
        _, sn, _, _ = flock_env.compute_helpers(g["x"][t], g["comm_radius"] ** 2)
        np.testing.assert_array_equal(eng.network_dense()[0], sn.astype(np.float32))
    if g["k"] > 1 and g["steps"] > 1:
        _, sn_prev, _, _ = flock_env.compute_helpers(g["x"][t - 1], g["comm_radius"] ** 2)
        np.testing.assert_array_equal(eng.network_dense(age=1)[0], sn_prev.astype(np.float32))
    eng.close()


def oracle_closed_loop(g, steps):
    """Closed-loop rollout of the numpy oracle from g['x'][0]."""
    layers = learner.weights_from_state_dict(g["state_dict"])
    x = g["x"][0].copy()
    R2 = g["comm_radius"] ** 2
    state, acts, rewards = None, [], []
    for t in range(steps):
        sv, sn, _, _ = flock_env.compute_helpers(x, R2)
        state = learner.DelayState((sv, sn), prev_state=state, k=g["k"])
        a = learner.select_action(layers, state)
        acts.append(a)
        x = flock_env.integrate(x, a, g["dt"])
        rewards.append(flock_env.instant_cost(x))
    return np.stack(acts), np.array(rewards), x


@pytest.mark.parametrize("name", ["ckpt_n100_k3", "rand_n100_k4_h64_l2", "rand_n64_k3_h128_l4", "rand_n50_k2_h16_l3"])
@pytest.mark.parametrize("readout", READOUTS)
def test_closed_loop_step_and_graph_rollout(name, readout):
    g = load_golden(name)
    if readout == TENSOR and g["hidden"] > 64:
        pytest.skip("tensor-core readout covers hidden <= 64")
    if readout == 3 and not 16 < g["hidden"] <= 64:
        pytest.skip("two-warp readout covers hidden in 17..64")
    T = 6
    acts_o, rew_o, x_o = oracle_closed_loop(g, T)
    # stepwise fused kernel
    eng = make_engine(g, readout_mode=readout)
    eng.reset(g["x"][0])
    acts, rews = [], []
    for t in range(T):
        a = np.empty((g["n_agents"], 2), np.float32)
        r = np.empty(1, np.float64)
        eng.step(a, r)
        acts.append(a)
        rews.append(r[0])
    acts = np.stack(acts)
    # free-running rollouts drift apart slowly (fp32 action differences feed back): looser bound
    assert rel_inf(acts, acts_o) <= 2e-4
    np.testing.assert_allclose(rews, rew_o, rtol=1e-5)
    x_step = eng.get_state()
    # CUDA-graph rollout must reproduce the stepwise path bit for bit
    eng2 = make_engine(g, readout_mode=readout)
    eng2.reset(g["x"][0])
    rew2 = eng2.rollout(T, want_reward=True)
    np.testing.assert_array_equal(eng2.get_state(), x_step)
    np.testing.assert_array_equal(eng2.get_action(), acts[-1])
    np.testing.assert_allclose(rew2[:, 0], rews, rtol=1e-12)
    # API-split path (policy + env_step) is the same arithmetic as the fused kernel
    eng3 = make_engine(g, readout_mode=readout)
    eng3.reset(g["x"][0])
    for t in range(T):
        a = eng3.policy().cpu().numpy()
        np.testing.assert_array_equal(a, acts[t])
        eng3.env_step(a)
    np.testing.assert_array_equal(eng3.get_state(), x_step)
    for e in (eng, eng2, eng3):
        e.close()


def test_run_to_run_reproducible():
    g = load_golden("ckpt_n400_k3_uniform")
    outs = []
    for _ in range(2):
        eng = make_engine(g)
        eng.reset(g["x"][0])
        eng.rollout(12)
        outs.append((eng.get_state(), eng.get_action(), eng.csr()))
        eng.close()
    np.testing.assert_array_equal(outs[0][0], outs[1][0])
    np.testing.assert_array_equal(outs[0][1], outs[1][1])
    for a, b in zip(outs[0][2][1:], outs[1][2][1:]):      # deg, cols, scale (row_start may differ)
        pass
    rs0, d0, c0, _ = outs[0][2]
    rs1, d1, c1, _ = outs[1][2]
    np.testing.assert_array_equal(d0, d1)
    for a in range(len(d0)):
        np.testing.assert_array_equal(c0[rs0[a]:rs0[a] + d0[a]], c1[rs1[a]:rs1[a] + d1[a]])


def test_batched_episodes_equal_separate_runs():
    g = load_golden("ckpt_n100_k3")
    from multiagent_gnn_policies_b200.engine import FlockEngine
    B, N, T = 5, g["n_agents"], 5
    rng = np.random.RandomState(0)
    xs = []
    for b in range(B):
        env = flock_env.FlockingRelativeOracle(n_agents=N, rng=rng)
        xs.append(env.sample_initial_state().copy())
    xs = np.stack(xs)
    # offsets in space must not matter, nor may episodes see each other even when they overlap
    big = FlockEngine(n_agents=N, n_episodes=B, k=3, hidden=32, n_layers=2)
    big.load_state_dict(g["state_dict"])
    big.reset(xs.reshape(B * N, 4))
    rew = big.rollout(T, want_reward=True)
    xb = big.get_state().reshape(B, N, 4)
    ab = big.get_action().reshape(B, N, 2)
    for b in range(B):
        one = make_engine(g)
        one.reset(xs[b])
        r1 = one.rollout(T, want_reward=True)
        np.testing.assert_array_equal(one.get_state(), xb[b])
        np.testing.assert_array_equal(one.get_action(), ab[b])
        np.testing.assert_allclose(r1[:, 0], rew[:, b], rtol=1e-12)
        one.close()
    big.close()


@pytest.mark.parametrize("n,R,hidden", [(20000, 1.0, 32), (30000, 2.0, 64)])
def test_large_n_against_sparse_oracle(n, R, hidden):
    """Sizes the dense reference cannot hold: compare with the edge-list oracle."""
    import torch
    from multiagent_gnn_policies_b200.engine import FlockEngine
    g = load_golden("ckpt_n100_k3")
    if hidden == 32:
        sd = g["state_dict"]
    else:
        torch.manual_seed(11)
        sd = {"conv_layers.0.weight": torch.randn(hidden, 6, 3, 1) * 0.2, "conv_layers.0.bias": torch.randn(hidden) * 0.1,
              "conv_layers.1.weight": torch.randn(hidden, hidden, 1, 1) * 0.1, "conv_layers.1.bias": torch.randn(hidden) * 0.1,
              "conv_layers.2.weight": torch.randn(2, hidden, 1, 1) * 0.1, "conv_layers.2.bias": torch.randn(2) * 0.1}
        sd = {k: v.numpy() for k, v in sd.items()}
    layers = learner.weights_from_state_dict(sd)
    x = flock_env.synthetic_state(n, seed=n, density=1.6)
    eng = FlockEngine(n_agents=n, k=3, hidden=hidden, n_layers=2, comm_radius=R, dt=0.01, edge_capacity=64)
    eng.load_state_dict(sd)
    eng.reset(x)
    sstate = None
    for t in range(4):
        sv, deg, i, j = sparse.compute_helpers_sparse(x, R)
        assert np.array_equal(eng.get_degrees(), deg)
        rs, dg, cols, scale = eng.csr()
        # identical edge SET per row (row order is the engine's cell order)
        got = np.concatenate([np.sort(cols[rs[a]:rs[a] + dg[a]]) for a in range(n)])
        assert np.array_equal(got, j)
        assert rel_inf(eng.get_features(), sv.astype(np.float32)) <= TOL_FEATURE
        a_net = sparse.network_csr(n, deg, i, j)
        sstate = sparse.SparseDelayState(sv, a_net, prev_state=sstate, k=3)
        act_o = sparse.readout(layers, sstate.aggregate())
        act = eng.policy().cpu().numpy()
        z = eng.get_aggregated()
        assert rel_inf(z, sstate.aggregate()) <= 1e-6
        # Among 2e4+ agents some sit where fp32 rounding of |z| ~ 1e3 inputs is amplified by the trained
        # weights (||W2|| ||W3|| ~ 40), so two correct fp32 evaluations differ by more than 1e-5 of ||a||.
        # Yardstick: the same inputs evaluated in float64.  The CUDA path must be as close to it as the
        # fp32 oracle (the reference arithmetic) is, or within the 1e-5 bar.
        truth = sparse.readout(layers, sstate.aggregate(np.float64), np.float64)
        log_parity(f"large_n n={n} h={hidden} t={t}", act, act_o, truth)
        assert rel_inf(act, truth) <= max(TOL_ACTION, 4.0 * rel_inf(act_o, truth))
        assert rel_inf(act, act_o) <= max(TOL_ACTION, 8.0 * rel_inf(act_o, truth))
        # drive the env with the expert (DAGGER with beta = 1): keeps agents apart, so |features| stay
        # O(1e3) and the fp32 readout stays well conditioned (a random policy makes agents collide)
        # ... with its float64 action, NOT cast to fp32: the reference steps the env with the controller's own array
        # (learner/gnn_dagger.py:156-163), the engine integrates it through fgnn_env_step_f64
        u = sparse.controller_sparse(x, R)
        assert u.dtype == np.float64
        x = flock_env.integrate(x, u, 0.01)
        eng.env_step(u)
        np.testing.assert_array_equal(eng.get_state(), x)
    assert not eng.stats()["overflow"]
    eng.close()


@pytest.mark.parametrize("name", ["ckpt_n100_k3", "rand_n100_k4_h64_l2", "rand_n64_k3_h128_l4", "rand_n12_k1_h4_l1"])
def test_dense_actor_forward_matches_reference(name):
    """Actor.forward(delay_state, delay_gso) on dense tensors, batch of consecutive states."""
    import torch
    g = load_golden(name)
    eng = make_engine(g)
    R2 = g["comm_radius"] ** 2
    state, ds, gso = None, [], []
    for t in range(g["steps"]):
        sv, sn, _, _ = flock_env.compute_helpers(g["x"][t], R2)
        state = learner.DelayState((sv, sn), prev_state=state, k=g["k"])
        ds.append(state.delay_state[0])
        gso.append(state.delay_gso[0])
    ds = torch.from_numpy(np.stack(ds)).cuda()
    gso = torch.from_numpy(np.stack(gso)).cuda()
    out = eng.actor_forward_dense(ds, gso).cpu().numpy()          # (B,1,2,N)
    assert out.shape == (g["steps"], 1, 2, g["n_agents"])
    ref = g["action"].transpose(0, 2, 1)[:, None]                 # (T,1,2,N)
    assert rel_inf(out, ref) <= TOL_ACTION
    eng.close()


def test_edge_cases():
    from multiagent_gnn_policies_b200.engine import FlockEngine, FgnnError
    g = load_golden("ckpt_n100_k3")
    layers = learner.weights_from_state_dict(g["state_dict"])
    # a single agent, and agents that never see each other: empty graph, action = MLP(0-features)
    for n, spread in ((1, 1.0), (7, 50.0)):
        eng = FlockEngine(n_agents=n, k=3, hidden=32, n_layers=2)
        eng.load_state_dict(g["state_dict"])
        x = np.zeros((n, 4))
        x[:, 0] = np.arange(n) * spread
        x[:, 2] = 1.0
        eng.reset(x)
        assert eng.get_degrees().sum() == 0 and eng.stats()["n_edges"] == 0
        a = eng.policy().cpu().numpy()
        exp = sparse.readout(layers, np.zeros((3, n, 6), np.float32))
        # |a| is only ~0.05 here (bias-only input): bound the absolute error at fp32-eps level instead
        assert np.abs(a - exp).max() <= 2e-6
        r = eng.env_step(a)
        assert r[0] == pytest.approx(flock_env.instant_cost(flock_env.integrate(x, a, 0.01)), abs=1e-12)
        eng.close()
    # negative coordinates / wrapped cell grid / N not a multiple of any block size
    n = 333
    x = flock_env.synthetic_state(n, seed=2, density=1.6)
    x[:, 0:2] -= 1000.25
    eng = FlockEngine(n_agents=n, k=3, hidden=32, n_layers=2, grid_dim=3)
    eng.load_state_dict(g["state_dict"])
    eng.reset(x)
    sv, sn, adj, deg = flock_env.compute_helpers(x, 1.0)
    assert np.array_equal(eng.get_degrees(), deg)
    np.testing.assert_array_equal(eng.network_dense()[0], sn.astype(np.float32))
    eng.close()
    # edge capacity exceeded -> sticky overflow flag, no crash
    eng = FlockEngine(n_agents=2000, k=3, hidden=32, n_layers=2, comm_radius=4.0, edge_capacity=1)
    eng.load_state_dict(g["state_dict"])
    eng.reset(flock_env.synthetic_state(2000, seed=4, density=1.6))
    assert eng.stats()["overflow"]
    eng.close()
    # bad configuration is rejected with a message
    with pytest.raises(FgnnError):
        FlockEngine(n_agents=10, k=9)


def _random_state_dict(rng, k, hidden, n_layers, scale=0.3):
    dims = [6] + [hidden] * n_layers + [2]
    sd = {}
    for i in range(len(dims) - 1):
        step = k if i == 0 else 1
        sd[f"conv_layers.{i}.weight"] = (rng.standard_normal((dims[i + 1], dims[i], step, 1)) * scale /
                                         np.sqrt(dims[i] * step)).astype(np.float32)
        sd[f"conv_layers.{i}.bias"] = (rng.standard_normal(dims[i + 1]) * 0.1).astype(np.float32)
    return sd


@pytest.mark.parametrize("k", [1, 2, 3, 4])
@pytest.mark.parametrize("hidden,n_layers", [(4, 1), (16, 2), (32, 3), (64, 4), (48, 2), (128, 2)])
@pytest.mark.parametrize("mean_pooling", [True, False])
def test_architecture_sweep_against_dense_oracle(k, hidden, n_layers, mean_pooling):
    """Every (K, H, L) the cfg sweeps use (cfg/hidden_size.cfg, cfg/n.cfg, k in 1..4), both pooling modes,
    both readouts: select_action checked step by step against the dense numpy oracle."""
    from multiagent_gnn_policies_b200.engine import FlockEngine
    rng = np.random.default_rng(1000 * k + hidden + n_layers)
    n = 150
    sd = _random_state_dict(rng, k, hidden, n_layers)
    layers = learner.weights_from_state_dict(sd)
    x0 = flock_env.synthetic_state(n, seed=k + hidden, density=1.6)
    for readout in (FFMA, TENSOR):
        if readout == TENSOR and hidden > 64:
            continue
        eng = FlockEngine(n_agents=n, k=k, hidden=hidden, n_layers=n_layers, comm_radius=1.2, dt=0.01,
                          mean_pooling=mean_pooling, readout_mode=readout, edge_capacity=64)
        eng.load_state_dict(sd)
        eng.reset(x0)
        x, state, sstate = x0.copy(), None, None
        for t in range(k + 3):
            sv, sn, _, deg = flock_env.compute_helpers(x, 1.2 ** 2, mean_pooling=mean_pooling)
            state = learner.DelayState((sv, sn), prev_state=state, k=k)
            a_ref = learner.select_action(layers, state)
            assert np.array_equal(eng.get_degrees(), deg)
            a = eng.policy().cpu().numpy()
            # yardstick: the same fp32 inputs evaluated in float64.  These random actors have |a| ~ 0.3 from
            # |features| ~ 1e2..1e3, so fp32 rounding alone is ~1e-5 of ||a||: the CUDA path must be as close to
            # the float64 value as the reference arithmetic (fp32 oracle) is, or within the 1e-5 bar.
            sv_s, deg_s, ei, ej = sparse.compute_helpers_sparse(x, 1.2)
            sstate = sparse.SparseDelayState(sv_s, sparse.network_csr(n, deg_s, ei, ej, mean_pooling), prev_state=sstate, k=k)
            truth = sparse.readout(layers, sstate.aggregate(np.float64), np.float64)
            log_parity(f"sweep k={k} h={hidden} l={n_layers} mp={int(mean_pooling)} readout={readout} t={t}", a, a_ref, truth)
            assert rel_inf(a, truth) <= max(TOL_ACTION, 3.0 * rel_inf(a_ref, truth)), (readout, t)
            assert rel_inf(a, a_ref) <= max(3.0 * TOL_ACTION, 6.0 * rel_inf(a_ref, truth)), (readout, t)
            # the expert drives (a random policy lets agents collide, |features| -> 1e6, fp32 ill-conditioned)
            u = flock_env.controller(x, 1.2, 1.2 ** 2, centralized=False)       # float64, as gym_flock returns it
            eng.env_step(u)
            x = flock_env.integrate(x, u, 0.01)
            np.testing.assert_array_equal(eng.get_state(), x)
        eng.close()


def test_full_size_one_million_agents():
    """BASELINE.json's headline size (N = 1M agents on one GPU, K = 3): exact edge set and degrees against the
    edge-list oracle, features to fp32 rounding, and the size-independent properties of the path -- a symmetric
    graph without self loops, antisymmetric pair terms that cancel over the flock, an integrator and a reward that
    equal the numpy expressions on the very actions the engine produced, run-to-run bit reproducibility."""
    from multiagent_gnn_policies_b200.engine import FlockEngine
    n, R = 1_000_000, 1.0
    g = load_golden("ckpt_n100_k3")
    x0 = flock_env.synthetic_state(n, seed=11, density=1.6)

    def new_engine():
        e = FlockEngine(n_agents=n, k=3, hidden=32, n_layers=2, comm_radius=R, dt=0.01, edge_capacity=32)
        e.load_state_dict(g["state_dict"])
        e.reset(x0)
        return e

    eng = new_engine()
    sv, deg, oi, oj = sparse.compute_helpers_sparse(x0, R)
    assert np.array_equal(eng.get_degrees(), deg)
    rs, dg, cols, _ = eng.csr()
    total = int(dg.sum())
    assert total == eng.stats()["n_edges"] == oi.size
    rows = np.repeat(np.arange(n, dtype=np.int64), dg)
    first = np.cumsum(dg, dtype=np.int64) - dg                              # position of each row's first edge in `rows`
    idx = np.repeat(rs.astype(np.int64), dg) + (np.arange(total, dtype=np.int64) - np.repeat(first, dg))
    c = cols[idx].astype(np.int64)
    assert np.all(rows != c)                                                # no self loops
    fwd = np.sort(rows * n + c)
    assert np.array_equal(fwd, np.sort(c * n + rows))                       # i in N(j)  <=>  j in N(i)
    assert np.array_equal(fwd, np.sort(oi.astype(np.int64) * n + oj.astype(np.int64)))      # the oracle's edge set
    feats = eng.get_features()
    assert rel_inf(feats, sv.astype(np.float32)) <= TOL_FEATURE
    f64 = feats.astype(np.float64)
    assert np.all(np.abs(f64.sum(axis=0)) <= 1e-6 * np.abs(f64).sum(axis=0))   # every pair term appears with both signs

    other = new_engine()
    x = x0
    a, a2 = np.empty((n, 2), np.float32), np.empty((n, 2), np.float32)
    r, r2 = np.empty(1, np.float64), np.empty(1, np.float64)
    for t in range(3):
        eng.step(a, r)
        other.step(a2, r2)
        assert np.array_equal(a, a2) and r[0] == r2[0]                       # bit-reproducible
        assert np.all(np.isfinite(a))
        x = flock_env.integrate(x, a, 0.01)
        np.testing.assert_array_equal(eng.get_state(), x)                   # the float64 integrator, bit for bit
        assert r[0] == pytest.approx(flock_env.instant_cost(x), rel=1e-9)
    np.testing.assert_array_equal(other.get_state(), x)
    assert not eng.stats()["overflow"] and not other.stats()["overflow"]
    eng.close()
    other.close()

    # aggregated features and actions at full size: three expert-driven steps (float64 controller action, as
    # learner/gnn_dagger.py:156-163 steps the env) against the edge-list oracle, like the 20k-agent test
    layers = learner.weights_from_state_dict(g["state_dict"])
    eng = new_engine()
    x, sstate = x0, None
    for t in range(3):
        sv, deg, oi, oj = sparse.compute_helpers_sparse(x, R)
        assert np.array_equal(eng.get_degrees(), deg)
        assert rel_inf(eng.get_features(), sv.astype(np.float32)) <= TOL_FEATURE
        sstate = sparse.SparseDelayState(sv, sparse.network_csr(n, deg, oi, oj), prev_state=sstate, k=3)
        act = eng.policy().cpu().numpy()
        assert rel_inf(eng.get_aggregated(), sstate.aggregate()) <= 2e-6
        act_o = sparse.readout(layers, sstate.aggregate())
        truth = sparse.readout(layers, sstate.aggregate(np.float64), np.float64)
        log_parity(f"one_million t={t}", act, act_o, truth)
        assert rel_inf(act, truth) <= max(TOL_ACTION, 4.0 * rel_inf(act_o, truth))
        assert rel_inf(act, act_o) <= max(TOL_ACTION, 8.0 * rel_inf(act_o, truth))
        u = sparse.controller_sparse(x, R)
        x = flock_env.integrate(x, u, 0.01)
        eng.env_step(u)
        np.testing.assert_array_equal(eng.get_state(), x)
    assert not eng.stats()["overflow"]
    eng.close()


@pytest.mark.parametrize("centralized", [False, True])
def test_float64_action_path(centralized):
    """env.step(env.controller()) (learner/gnn_dagger.py:156-163, learner/gnn_baseline.py:16-17): the controller's action
    is a float64 array and the reference steps the env with it as it is.  The engine's float64 entry points
    (fgnn_controller_f64 -> fgnn_env_step_f64) must (a) integrate a float64 action bit for bit like the oracle -- the
    oracle's action is NOT rounded to fp32 first -- and (b) produce a controller action equal to the oracle's to float64
    accuracy, so that the closed expert loop stays on the oracle's trajectory far below fp32 resolution."""
    from multiagent_gnn_policies_b200.engine import FlockEngine
    n, R = 400, 1.0
    x = flock_env.synthetic_state(n, seed=21, density=1.6)
    eng = FlockEngine(n_agents=n, comm_radius=R, edge_capacity=64)
    eng.reset(x)
    xe = x.copy()                                           # trajectory driven by the ENGINE's float64 controller
    for t in range(5):
        u_ref = flock_env.controller(x, R, R * R, centralized=centralized)
        assert u_ref.dtype == np.float64 and np.any(u_ref != u_ref.astype(np.float32))     # genuinely float64 values
        u = eng.controller(centralized=centralized, dtype=np.float64)
        assert u.dtype == np.float64
        np.testing.assert_allclose(u, u_ref, rtol=1e-9, atol=1e-11)
        # (a) teacher-forced: the oracle's own float64 action through the engine's integrator
        eng.env_step(u_ref)
        x = flock_env.integrate(x, u_ref, 0.01)
        np.testing.assert_array_equal(eng.get_state(), x)
        # the fp32 entry point on the same action differs (that is what the old path did)
        assert not np.array_equal(flock_env.integrate(xe, u_ref.astype(np.float32), 0.01), flock_env.integrate(xe, u_ref, 0.01))
        xe = x
    # torch float64 CUDA tensors take the same path
    import torch
    u_ref = flock_env.controller(x, R, R * R, centralized=centralized)
    eng.env_step(torch.from_numpy(u_ref).cuda())
    np.testing.assert_array_equal(eng.get_state(), flock_env.integrate(x, u_ref, 0.01))
    eng.close()


@pytest.mark.parametrize("n,episodes,k,radius,tail_only", [
    (100, 1, 3, 1.0, False),          # 10 x 10 grid: every warp straddles grid rows (per-lane path)
    (200000, 1, 3, 1.0, False),       # 447-cell rows: nearly every warp takes the staged path
    (150000, 1, 3, 1.0, True),
    (3000, 1, 3, 1.0, True),
    (20000, 1, 4, 1.0, False),
    (5000, 1, 2, 1.0, False),
    (4000, 1, 1, 1.0, False),
    (500, 6, 3, 1.0, False),          # batched episodes
    (20000, 3, 3, 1.0, False),        # batched episodes with grid rows long enough for the staged path (episode > 0)
    (3000, 1, 3, 3.0, False),         # ~45 neighbours: cell rows longer than 32 candidates, rows longer than the staged
                                      # neighbour list, warps that outgrow the stage
    (2500, 1, 3, 2.0, True),
])
def test_pair_kernels_equal_separate_kernels(n, episodes, k, radius, tail_only, monkeypatch):
    """k_pair_filter + k_pair_fused (FGNN_STEP_MODE=1: adjacency + features + first hop from warp-staged row ranges, TMA bulk
    staging, fp32 pre-filter with the float64 test for pairs inside the margin) must leave the same bits as k_adjacency_t +
    k_hop (FGNN_STEP_MODE=0): degrees, features, aggregated z, actions and the integrated state, over a closed-loop rollout."""
    from multiagent_gnn_policies_b200.engine import FlockEngine
    rng = np.random.default_rng(n + k)
    sd = _random_state_dict(rng, k, 32, 2)
    x0 = np.concatenate([flock_env.synthetic_state(n, seed=7 + e, density=1.6) for e in range(episodes)])
    cap = int(max(48, 3.5 * np.pi * radius ** 2 * 1.6 + 16))
    runs = []
    for mode in ("0", "1"):
        monkeypatch.setenv("FGNN_STEP_MODE", mode)
        eng = FlockEngine(n_agents=n, n_episodes=episodes, k=k, hidden=32, n_layers=2, comm_radius=radius, dt=0.01,
                          edge_capacity=cap, csr_tail_only=(tail_only and mode == "1"))
        eng.load_state_dict(sd)
        eng.reset(x0)
        out = []
        for t in range(8):
            a = np.empty((n * episodes, 2), np.float32)
            d, f = eng.get_degrees(), eng.get_features()
            eng.policy(out=a)
            out.append((d, f, eng.get_aggregated(), a.copy(), eng.get_state()))
            eng.env_step(a * 0.05)        # a tame closed loop (random weights): agents keep moving across cells
        eng.rollout(6)                    # ... and the CUDA-graph replay of the same kernels
        out.append((eng.get_degrees(), eng.get_features(), eng.get_state()))
        assert not eng.stats()["overflow"]
        # complete CSR rows (pair mode: assembled on demand from the ELL head + the tail of long rows): same neighbour order
        rs, dg, cols, _ = eng.csr()
        runs.append((out, np.concatenate([cols[rs[a_]:rs[a_] + dg[a_]] for a_ in range(min(n * episodes, 2000))])))
        eng.close()
    (a_out, a_cols), (b_out, b_cols) = runs
    for t, (ra, rb) in enumerate(zip(a_out, b_out)):
        for u, v in zip(ra, rb):
            np.testing.assert_array_equal(u, v, err_msg=f"step {t}")
    np.testing.assert_array_equal(a_cols, b_cols)


@pytest.mark.parametrize("n,k,hidden", [(300001, 3, 32), (270000, 2, 16), (262144, 4, 64), (280000, 1, 32), (300001, 3, 128)])
def test_host_buffer_policy_chunked_equals_single_launch(n, k, hidden, monkeypatch):
    """fgnn_policy into a HOST buffer (select_action -> numpy, the reference-facing call) at sizes that take the chunked
    path (M >= 2^18): the pipelined form (FGNN_POLICY_PIPE: last hop inside every chunk's readout, copies on a second stream)
    and the equal-chunk form must fill the same actions / aggregated z as ONE readout launch into device memory, and the
    step that follows must leave the same state."""
    import torch
    from multiagent_gnn_policies_b200.engine import FlockEngine
    rng = np.random.default_rng(n + k)
    sd = _random_state_dict(rng, k, hidden, 2)
    x0 = flock_env.synthetic_state(n, seed=3, density=1.6)
    runs = []
    for chunks, pipe, host in ((1, 0, False), (4, 0, True), (4, 1, True), (4, 2, True), (4, 5, True)):
        monkeypatch.setenv("FGNN_POLICY_CHUNKS", str(chunks))
        monkeypatch.setenv("FGNN_POLICY_PIPE", str(pipe))
        eng = FlockEngine(n_agents=n, k=k, hidden=hidden, n_layers=2, comm_radius=1.0, dt=0.01)
        eng.load_state_dict(sd)
        eng.reset(x0)
        eng.rollout(k + 1)
        out = []
        for t in range(3):
            if host:
                a = torch.full((n, 2), float("nan"), dtype=torch.float32).pin_memory().numpy()
                eng.policy(out=a)
            else:
                a = eng.policy().cpu().numpy()
            assert np.isfinite(a).all()
            out += [a.copy(), eng.get_aggregated(), eng.get_action()]
            eng.env_step(a * 0.05)
            out.append(eng.get_state())
        assert not eng.stats()["overflow"]
        runs.append(out)
        eng.close()
    for r in runs[1:]:
        for u, v in zip(runs[0], r):
            np.testing.assert_array_equal(u, v)
