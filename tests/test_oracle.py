"""The oracle against the reference-generated golden vectors, and the sparse oracle
against the dense one (CPU only)."""
import numpy as np
import pytest

from conftest import rel_inf, load_golden, agg_golden_names, load_agg_golden
from oracle import flock_env, learner, sparse

TOL_ACTION = 1e-5      # north_star: actions within 1e-5 relative fp32


def replay(g):
    """Teacher-forced replay of a golden trajectory through the numpy oracle."""
    layers = learner.weights_from_state_dict(g["state_dict"])
    R2 = g["comm_radius"] ** 2
    state = None
    for t in range(g["steps"]):
        sv, sn, adj, deg = flock_env.compute_helpers(g["x"][t], R2)
        state = learner.DelayState((sv, sn), prev_state=state, k=g["k"])
        yield t, sv, deg, state, layers


def test_env_restatement_matches_recorded_trajectory(golden):
    g = golden
    for t, sv, deg, state, layers in replay(g):
        assert np.array_equal(deg, g["deg"][t])
        np.testing.assert_array_equal(sv, g["values"][t])
        if t + 1 < g["steps"]:
            xn = flock_env.integrate(g["x"][t], g["action"][t], g["dt"])
            np.testing.assert_array_equal(xn, g["x"][t + 1])
            assert flock_env.instant_cost(xn) == pytest.approx(float(g["reward"][t]), rel=1e-12)


def test_learner_restatement_matches_reference(golden):
    g = golden
    for t, sv, deg, state, layers in replay(g):
        z = learner.aggregate(state.delay_state, state.delay_gso)[0]
        assert rel_inf(z, g["z"][t]) <= 2e-6
        a = learner.select_action(layers, state)
        assert rel_inf(a, g["action"][t]) <= TOL_ACTION


@pytest.mark.parametrize("name", agg_golden_names())
def test_actor_restatement_for_any_aggregation_index(name):
    """oracle.learner.actor_forward_any against the UNMODIFIED reference Actor.forward with ind_agg in {0, 1, 2} and unequal
    layer widths (oracle/gen_golden_agg.py): fp32 sums in a different order, so 2e-6 of the output's inf-norm."""
    g = load_agg_golden(name)
    layers = learner.weights_from_state_dict(g["state_dict"])
    out = learner.actor_forward(layers, g["delay_state"], g["delay_gso"], ind_agg=g["ind_agg"])
    assert out.shape == g["out"].shape
    assert rel_inf(out, g["out"]) <= 2e-6


def test_delay_state_structure():
    g = load_golden("ckpt_n100_k3")
    states = [s for _, _, _, s, _ in replay(g)]
    s0, s1, s2 = states[0], states[1], states[2]
    n = g["n_agents"]
    assert np.array_equal(s0.delay_gso[0, 0], np.eye(n, dtype=np.float32))
    assert not s0.delay_gso[0, 1:].any() and not s0.delay_state[0, 1:].any()
    assert np.array_equal(s1.delay_gso[0, 1], s1.network[0, 0])
    assert not s1.delay_gso[0, 2].any()
    np.testing.assert_allclose(s2.delay_gso[0, 2], s2.network[0, 0] @ s1.network[0, 0], rtol=1e-6, atol=1e-8)
    assert np.array_equal(s2.delay_state[0, 2], s0.delay_state[0, 0])


@pytest.mark.parametrize("n,R", [(150, 1.0), (400, 1.5), (64, 0.8)])
def test_sparse_oracle_equals_dense(n, R):
    x = flock_env.synthetic_state(n, seed=n, density=1.6)
    sv, sn, adj, deg = flock_env.compute_helpers(x, R * R)
    sv_s, deg_s, i, j = sparse.compute_helpers_sparse(x, R)
    assert np.array_equal(deg, deg_s)
    ii, jj = np.nonzero(adj)
    assert np.array_equal(ii, i) and np.array_equal(jj, j)
    np.testing.assert_allclose(sv_s, sv, rtol=1e-12, atol=1e-9)
    a = sparse.network_csr(n, deg_s, i, j).toarray()
    np.testing.assert_array_equal(a, sn.astype(np.float32))


def test_sparse_controller_equals_dense():
    x = flock_env.synthetic_state(200, seed=9, density=1.6)
    for R in (1.0, 1.7):
        u_d = flock_env.controller(x, R, R * R, centralized=False)
        u_s = sparse.controller_sparse(x, R)
        np.testing.assert_allclose(u_s, u_d, rtol=1e-10, atol=1e-12)


def test_sparse_delay_state_equals_dense_path():
    g = load_golden("rand_n100_k4_h64_l2")
    layers = learner.weights_from_state_dict(g["state_dict"])
    R = g["comm_radius"]
    sstate = None
    for t in range(g["steps"]):
        sv, deg, i, j = sparse.compute_helpers_sparse(g["x"][t], R)
        a = sparse.network_csr(g["n_agents"], deg, i, j)
        sstate = sparse.SparseDelayState(sv, a, prev_state=sstate, k=g["k"])
        z = sstate.aggregate()                       # (K,N,F)
        assert rel_inf(z.transpose(0, 2, 1), g["z"][t]) <= 2e-6
        act = sparse.readout(layers, z)
        assert rel_inf(act, g["action"][t]) <= TOL_ACTION


def test_integrator_and_cost_small_case():
    x = np.array([[0.0, 0.0, 1.0, 0.0], [1.0, 2.0, 0.0, -1.0]])
    u = np.array([[0.1, 0.0], [0.0, -0.2]])
    xn = flock_env.integrate(x, u, 0.5, action_scalar=10.0)
    np.testing.assert_allclose(xn, [[0.5 + 0.125, 0.0, 1.5, 0.0], [1.0, 2.0 - 0.5 - 0.25, 0.0, -2.0]])
    assert flock_env.instant_cost(xn) == pytest.approx(-(np.var([1.5, 0.0]) + np.var([0.0, -2.0])))


def test_controller_shapes_and_clip():
    env = flock_env.FlockingRelativeOracle(n_agents=30, rng=np.random.RandomState(0))
    env.reset()
    for c in (True, False, None):
        u = env.controller(c)
        assert u.shape == (30, 2) and np.all(np.abs(u) <= 1.0 + 1e-12)


# ---- environment variants (parity unpinned like the base env) ----------------------

def test_leader_variant_masks_the_action():
    rng = np.random.RandomState(4)
    env = flock_env.FlockingLeaderOracle(n_agents=40, rng=rng)
    env.reset()
    v_lead = env.x[0:2, 2:4].copy()
    assert np.all(v_lead == v_lead[0, 0])                      # one shared scalar velocity
    x0 = env.x.copy()
    u = rng.uniform(-1, 1, size=(40, 2))
    env.step(u)
    np.testing.assert_array_equal(env.x[0:2, 2:4], v_lead)     # leaders keep their velocity exactly
    np.testing.assert_array_equal(env.x[0:2, 0:2], x0[0:2, 0:2] + v_lead * env.dt)
    free = flock_env.integrate(x0, u, env.dt)
    np.testing.assert_array_equal(env.x[2:], free[2:])         # followers: the plain double integrator


def test_two_flocks_variant_initial_state():
    env = flock_env.FlockingTwoFlocksOracle(n_agents=100, rng=np.random.RandomState(2))
    env.reset()
    half = 50
    assert env.x[:half, 0].mean() < 0 < env.x[half:, 0].mean()
    d = env.x[:half, 2:4].mean(axis=0) + env.x[half:, 2:4].mean(axis=0)     # opposite biases cancel
    assert np.abs(d).max() < 4 * env.v_max / np.sqrt(half)
    sv, sn = env.helpers()
    assert (np.asarray(sn) != 0).sum(axis=1).min() >= 2


def test_stochastic_variant_draws_dt_per_step():
    env = flock_env.FlockingStochasticOracle(n_agents=30, v_max=0.5, comm_radius=1.5, rng=np.random.RandomState(3))
    env.reset()
    dts = []
    for _ in range(5):
        x0 = env.x.copy()
        u = env.controller(False)
        env.step(u)
        dts.append(env.dt)
        np.testing.assert_array_equal(env.x, flock_env.integrate(x0, u, env.dt))
    assert len(set(dts)) == 5 and min(dts) >= env.dt_min


def test_env_oracle_against_real_gym_flock():
    """Pin of the env oracle against the real package (github.com/katetolstaya/gym-flock, un-vendored and un-pinned by the
    reference: README.md:7, train.py:6).  gym_flock is not installed in this image, so this test SKIPS here and the env
    side stays `parity unpinned`; wherever the package is importable it compares, on the package's own reset state:
    state_values / state_network of compute_helpers, the double-integrator step, the reward and both controllers."""
    import sys
    # the REAL package only: the compat shim of the same name (multiagent_gnn_policies_b200/compat, possibly put on
    # sys.path by an earlier test) is the thing under test elsewhere, not a pin
    shim = [m for m in list(sys.modules) if m == "gym_flock" or m.startswith("gym_flock.")]
    saved_modules = {m: sys.modules.pop(m) for m in shim}
    saved_path = list(sys.path)
    sys.path[:] = [q for q in sys.path if "multiagent_gnn_policies_b200" not in q]
    try:
        gym_flock = pytest.importorskip("gym_flock")
    finally:
        sys.path[:] = saved_path
        if "gym_flock" not in sys.modules:
            sys.modules.update(saved_modules)
    import configparser
    cp = configparser.ConfigParser()
    cp.read_dict({"DEFAULT": {"comm_radius": "1.0", "n_agents": "60", "v_max": "3.0", "dt": "0.01"}})
    env = gym_flock.envs.FlockingRelativeEnv()
    env.params_from_cfg(cp["DEFAULT"])
    np.random.seed(7)
    values, network = env.reset()
    x = np.array(env.x, dtype=np.float64)
    R2 = float(env.comm_radius2)
    for t in range(5):
        sv, sn, _, _ = flock_env.compute_helpers(x, R2, mean_pooling=bool(getattr(env, "mean_pooling", True)))
        np.testing.assert_allclose(values, sv, rtol=1e-12, atol=1e-12)
        np.testing.assert_array_equal(np.asarray(network) != 0, sn != 0)
        np.testing.assert_allclose(network, sn, rtol=1e-15, atol=0)
        for centralized in (True, False):
            u_env = np.asarray(env.controller(centralized))
            u_or = flock_env.controller(x, float(env.comm_radius), R2, centralized=centralized,
                                        max_accel=float(getattr(env, "max_accel", 1.0)),
                                        action_scalar=float(getattr(env, "action_scalar", 10.0)))
            np.testing.assert_allclose(u_env, u_or, rtol=1e-10, atol=1e-12)
        u = np.asarray(env.controller(False))
        (values, network), reward, done, _ = env.step(u)
        x = flock_env.integrate(x, u, float(env.dt), action_scalar=float(getattr(env, "action_scalar", 10.0)))
        np.testing.assert_array_equal(np.asarray(env.x, dtype=np.float64), x)
        assert reward == pytest.approx(flock_env.instant_cost(x), rel=1e-12)
        assert done is False or done == 0


def test_reference_staging_is_byte_identical():
    """oracle/_ref (bench.py's reference leg, tests/test_gpu_reference_scripts.py) holds byte-for-byte copies: the sha256
    of every staged file equals the manifest written when it was copied from the reference tree; when that tree is here
    the copies are compared with it directly."""
    import filecmp
    import os
    from oracle import make_ref
    if not make_ref.available():
        pytest.skip("oracle/_ref not staged (run `python -m oracle.make_ref` where the reference tree is present)")
    assert make_ref.verify()
    if os.path.isdir(os.path.join(make_ref.REF, "learner")):
        for rel in make_ref.FILES:
            assert filecmp.cmp(os.path.join(make_ref.REF, rel), os.path.join(make_ref.DST, rel), shallow=False), rel
