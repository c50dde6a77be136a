"""numpy float64 restatement of gym_flock's ``FlockingRelativeEnv``.

TEST INFRASTRUCTURE (see oracle/__init__.py).  **parity unpinned**: gym_flock is
an un-vendored, un-pinned dependency of the reference (README.md:7, train.py:6);
this file restates its published algorithm.  The only
contracts the reference itself pins are at learner/state_with_delay.py:22-26:
``state_values.shape == (N, n_states)``, ``state_network.shape == (N, N)`` and a
zero diagonal.  Every constant is a parameter.

All routines are dense O(N^2) like the upstream code, so they are for small N;
``oracle.sparse`` holds the same arithmetic on an edge list for large N.
"""
import numpy as np

N_FEATURES = 6
NX = 4          # px, py, vx, vy
NU = 2


def pair_terms(x):
    """diff[i,j,:] = x_i - x_j and r2[i,j] = dx*dx + dy*dy with an infinite diagonal."""
    n = x.shape[0]
    diff = x.reshape(n, 1, NX) - x.reshape(1, n, NX)
    r2 = np.multiply(diff[:, :, 0], diff[:, :, 0]) + np.multiply(diff[:, :, 1], diff[:, :, 1])
    np.fill_diagonal(r2, np.inf)
    return diff, r2


def compute_helpers(x, comm_radius2, mean_pooling=True):
    """Graph + features from the 4-d agent state (SURVEY Appendix B ``compute_helpers``).

    Returns (state_values (N,6) f64, state_network (N,N) f64, adj (N,N) f64, deg (N,) int).
    Feature order: [dvx, dpx/r^4, dpx/r^2, dvy, dpy/r^4, dpy/r^2], summed over radius neighbours.
    """
    n = x.shape[0]
    diff, r2 = pair_terms(x)
    adj = (r2 < comm_radius2).astype(np.float64)
    deg = adj.sum(axis=1)
    nn = deg.reshape(n, 1).copy()
    nn[nn == 0] = 1
    adj_mean = adj / nn
    with np.errstate(divide="ignore", invalid="ignore"):
        r4 = np.multiply(r2, r2)
        feats = np.dstack((diff[:, :, 2], diff[:, :, 0] / r4, diff[:, :, 0] / r2,
                           diff[:, :, 3], diff[:, :, 1] / r4, diff[:, :, 1] / r2))
    feats[~np.isfinite(feats)] = 0.0          # diagonal (r2 = inf) terms
    state_values = np.sum(feats * adj.reshape(n, n, 1), axis=1).reshape(n, N_FEATURES)
    state_network = adj_mean if mean_pooling else adj
    return state_values, state_network, adj, deg.astype(np.int64)


def integrate(x, u, dt, action_scalar=10.0, half_accel_term=True, mask=None):
    """Double integrator (Appendix B ``step``): a = u*gain; p += v dt (+ a dt^2/2); v += a dt.
    ``mask`` (N,) of 0/1: FlockingLeader's ``u * mask`` -- agents with mask 0 ignore their action."""
    x = x.copy()
    a = np.asarray(u, dtype=np.float64) * action_scalar
    if mask is not None:
        a = a * np.asarray(mask, dtype=np.float64).reshape(-1, 1)
    if half_accel_term:
        x[:, 0] = x[:, 0] + x[:, 2] * dt + a[:, 0] * dt * dt * 0.5
        x[:, 1] = x[:, 1] + x[:, 3] * dt + a[:, 1] * dt * dt * 0.5
    else:
        x[:, 0] = x[:, 0] + x[:, 2] * dt
        x[:, 1] = x[:, 1] + x[:, 3] * dt
    x[:, 2] = x[:, 2] + a[:, 0] * dt
    x[:, 3] = x[:, 3] + a[:, 1] * dt
    return x


def instant_cost(x):
    """Reward returned by ``step``: minus the summed per-axis velocity variance."""
    return -1.0 * float(np.sum(np.var(x[:, 2:4], axis=0)))


def potential_grad(d, r2):
    return -2.0 * d / (r2 * r2) + 2.0 * d / r2


def controller(x, comm_radius, comm_radius2, centralized=True, max_accel=1.0, action_scalar=10.0):
    """Expert potential-based controller (Appendix B ``controller``), (N,2) in action units."""
    diff, r2 = pair_terms(x)
    adj = (r2 < comm_radius2).astype(np.float64)
    with np.errstate(divide="ignore", invalid="ignore"):
        gx = potential_grad(diff[:, :, 0], r2)
        gy = potential_grad(diff[:, :, 1], r2)
    gx[~np.isfinite(gx)] = 0.0
    gy[~np.isfinite(gy)] = 0.0
    far = r2 > comm_radius          # upstream compares r2 against R (not R^2); kept as recorded
    gx[far] = 0.0
    gy[far] = 0.0
    pot = np.dstack((diff[:, :, 2], diff[:, :, 3], gx, gy))
    if not centralized:
        pot = pot * adj.reshape(adj.shape[0], adj.shape[1], 1)
    p = pot.sum(axis=1)
    u = np.stack((-p[:, 2] - p[:, 0], -p[:, 3] - p[:, 1]), axis=1)
    lim = max_accel * action_scalar
    return np.clip(u, -lim, lim) / action_scalar


class FlockingRelativeOracle:
    """Stateful wrapper with the env call surface the reference uses
    (reset/step/controller/params_from_cfg; train.py:17-25, gnn_dagger.py:150-163)."""

    def __init__(self, n_agents=100, comm_radius=1.0, v_max=3.0, dt=0.01, r_max0=1.0,
                 action_scalar=10.0, max_accel=1.0, mean_pooling=True, half_accel_term=True,
                 min_dist_thresh=0.1, min_degree=2, rng=None):
        self.n_agents = n_agents
        self.comm_radius = comm_radius
        self.comm_radius2 = comm_radius * comm_radius
        self.v_max = v_max
        self.v_bias = v_max
        self.dt = dt
        self.r_max0 = r_max0
        self.r_max = r_max0 * np.sqrt(n_agents)
        self.action_scalar = action_scalar
        self.max_accel = max_accel
        self.mean_pooling = mean_pooling
        self.half_accel_term = half_accel_term
        self.min_dist_thresh = min_dist_thresh
        self.min_degree = min_degree
        self.rng = rng if rng is not None else np.random
        self.x = np.zeros((n_agents, NX))

    def params_from_cfg(self, args):
        self.comm_radius = args.getfloat('comm_radius')
        self.comm_radius2 = self.comm_radius * self.comm_radius
        self.n_agents = args.getint('n_agents')
        self.r_max = self.r_max0 * np.sqrt(self.n_agents)
        self.v_max = args.getfloat('v_max')
        self.v_bias = self.v_max
        self.dt = args.getfloat('dt')

    def sample_initial_state(self, max_tries=100000):
        n = self.n_agents
        x = np.zeros((n, NX))
        for _ in range(max_tries):
            length = np.sqrt(self.rng.uniform(0, self.r_max, size=(n,)))
            angle = np.pi * self.rng.uniform(0, 2, size=(n,))
            x[:, 0] = length * np.cos(angle)
            x[:, 1] = length * np.sin(angle)
            bias = self.rng.uniform(low=-self.v_bias, high=self.v_bias, size=(2,))
            x[:, 2] = self.rng.uniform(low=-self.v_max, high=self.v_max, size=(n,)) + bias[0]
            x[:, 3] = self.rng.uniform(low=-self.v_max, high=self.v_max, size=(n,)) + bias[1]
            _, r2 = pair_terms(x)
            min_dist = np.sqrt(np.min(r2))
            degree = np.min(np.sum((r2 < self.comm_radius2).astype(int), axis=1))
            if degree >= self.min_degree and min_dist >= self.min_dist_thresh:
                return x
        raise RuntimeError("no admissible initial configuration found")

    def helpers(self):
        sv, sn, _, _ = compute_helpers(self.x, self.comm_radius2, self.mean_pooling)
        return sv, sn

    def reset(self, x0=None):
        self.x = self.sample_initial_state() if x0 is None else np.array(x0, dtype=np.float64)
        return self.helpers()

    def step(self, u):
        u = np.asarray(u)
        assert u.shape == (self.n_agents, NU)
        self.x = integrate(self.x, u, self.dt, self.action_scalar, self.half_accel_term)
        return self.helpers(), instant_cost(self.x), False, {}

    def controller(self, centralized=None):
        if centralized is None:
            centralized = True
        return controller(self.x, self.comm_radius, self.comm_radius2, centralized,
                          self.max_accel, self.action_scalar)


class FlockingLeaderOracle(FlockingRelativeOracle):
    """FlockingLeader-v0 (cfg/dagger_leader.cfg:24, cfg/vel_leader_baseline.cfg) [UNVERIFIED-MEMORY of gym_flock]:
    the first ``n_leaders`` agents share one constant velocity and ignore every action (``u * mask``).
    The observation returned by reset() is recomputed after the leaders' velocities are set."""

    def __init__(self, n_leaders=2, **kw):
        super().__init__(**kw)
        self.n_leaders = n_leaders
        self.mask = np.ones((self.n_agents,))
        self.mask[0:n_leaders] = 0

    def params_from_cfg(self, args):
        super().params_from_cfg(args)
        self.mask = np.ones((self.n_agents,))
        self.mask[0:self.n_leaders] = 0

    def reset(self, x0=None):
        super().reset(x0)
        if x0 is None:
            self.x[0:self.n_leaders, 2:4] = np.ones((self.n_leaders, 2)) * self.rng.uniform(
                low=-self.v_max, high=self.v_max, size=(1, 1))
        return self.helpers()

    def step(self, u):
        u = np.asarray(u)
        assert u.shape == (self.n_agents, NU)
        self.x = integrate(self.x, u, self.dt, self.action_scalar, self.half_accel_term, mask=self.mask)
        return self.helpers(), instant_cost(self.x), False, {}


class FlockingTwoFlocksOracle(FlockingRelativeOracle):
    """FlockingTwoFlocks-v0 (cfg/dagger_twoflocks.cfg:24, cfg/n_twoflocks.cfg) [UNVERIFIED-MEMORY of gym_flock]:
    same dynamics; reset() draws two half-flocks -- discs of half the area each, so the density equals the single
    flock's -- whose centres are ``flock_offset`` apart along x (default: tangent discs), with opposite velocity
    biases so that they fly through each other."""

    def __init__(self, flock_offset=None, **kw):
        super().__init__(**kw)
        self.flock_offset = flock_offset

    def sample_initial_state(self, max_tries=100000):
        n = self.n_agents
        half = n // 2
        x = np.zeros((n, NX))
        offset = 2.0 * np.sqrt(0.5 * self.r_max) if self.flock_offset is None else self.flock_offset
        for _ in range(max_tries):
            length = np.sqrt(self.rng.uniform(0, 0.5 * self.r_max, size=(n,)))
            angle = np.pi * self.rng.uniform(0, 2, size=(n,))
            x[:, 0] = length * np.cos(angle)
            x[:, 1] = length * np.sin(angle)
            x[:half, 0] -= 0.5 * offset
            x[half:, 0] += 0.5 * offset
            bias = self.rng.uniform(low=-self.v_bias, high=self.v_bias, size=(2,))
            sign = np.where(np.arange(n) < half, 1.0, -1.0)
            x[:, 2] = self.rng.uniform(low=-self.v_max, high=self.v_max, size=(n,)) + sign * bias[0]
            x[:, 3] = self.rng.uniform(low=-self.v_max, high=self.v_max, size=(n,)) + sign * bias[1]
            _, r2 = pair_terms(x)
            min_dist = np.sqrt(np.min(r2))
            degree = np.min(np.sum((r2 < self.comm_radius2).astype(int), axis=1))
            if degree >= self.min_degree and min_dist >= self.min_dist_thresh:
                return x
        raise RuntimeError("no admissible initial configuration found")


class FlockingStochasticOracle(FlockingRelativeOracle):
    """FlockingStochastic-v0 (cfg/dagger_stoch.cfg:24, cfg/rad_stoch.cfg, cfg/transfer_stoch.cfg -- none of which
    carries a ``dt`` key) [UNVERIFIED-MEMORY of gym_flock]: the time step of every env.step is random,
    dt ~ max(N(dt_mean, dt_sigma), dt_min), imitating a simulator with an irregular clock."""

    def __init__(self, dt_mean=0.1, dt_sigma=0.02, dt_min=1e-3, **kw):
        kw.setdefault("dt", dt_mean)
        super().__init__(**kw)
        self.dt_mean, self.dt_sigma, self.dt_min = dt_mean, dt_sigma, dt_min

    def params_from_cfg(self, args):
        dt = self.dt
        super().params_from_cfg(args)
        if self.dt is None:
            self.dt = dt

    def draw_dt(self):
        return float(max(self.rng.normal(self.dt_mean, self.dt_sigma), self.dt_min))

    def step(self, u):
        self.dt = self.draw_dt()
        return super().step(u)


def synthetic_state(n_agents, seed=11, density=1.6, v_max=3.0, sort_cells=True, cell=1.0, min_dist=0.1,
                    dtype=np.float64):
    """Benchmark workload of SURVEY.md section 8(d): uniform positions in a square of side
    sqrt(N/density), velocities U(-v_max,v_max)+bias, optionally in cell-major order.
    Like gym_flock's reset (min_dist_thresh = 0.1) no two agents start closer than ``min_dist``:
    offending points are re-drawn (near-coincident pairs make dp/r^4 ~ 1e9 and the fp32 readout
    ill-conditioned for every implementation)."""
    rng = np.random.default_rng(seed)
    side = np.sqrt(n_agents / density)
    x = np.empty((n_agents, NX), dtype=np.float64)
    x[:, 0:2] = rng.uniform(0.0, side, size=(n_agents, 2))
    if min_dist > 0 and n_agents > 1:
        from scipy.spatial import cKDTree
        for _ in range(200):
            pairs = cKDTree(x[:, 0:2]).query_pairs(min_dist, output_type='ndarray')
            if pairs.size == 0:
                break
            bad = np.unique(pairs[:, 1])
            x[bad, 0:2] = rng.uniform(0.0, side, size=(bad.size, 2))
        else:
            raise RuntimeError("could not separate agents by min_dist")
    bias = rng.uniform(-v_max, v_max, size=(2,))
    x[:, 2:4] = rng.uniform(-v_max, v_max, size=(n_agents, 2)) + bias
    if sort_cells:
        cx = np.floor(x[:, 0] / cell).astype(np.int64)
        cy = np.floor(x[:, 1] / cell).astype(np.int64)
        order = np.lexsort((cx, cy))
        x = x[order]
    return x.astype(dtype)
