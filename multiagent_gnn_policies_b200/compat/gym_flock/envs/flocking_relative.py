"""FlockingRelativeEnv on the CUDA engine.

Call surface used by the reference: ``params_from_cfg`` (train.py:21), ``seed`` (train.py:25),
``reset``/``step`` returning ``(state_values (N,6), state_network (N,N))`` (learner/state_with_delay.py:22-26),
``controller(centralized=None)`` (learner/gnn_dagger.py:156, learner/gnn_baseline.py:16), ``render``, ``close``.
Arithmetic follows SURVEY.md Appendix B (gym_flock itself is absent: parity unpinned for the env).
All per-step arithmetic runs in libfgnn.so; the host only samples the initial configuration.
"""
import numpy as np

from multiagent_gnn_policies_b200.engine import FlockEngine, FgnnError

DENSE_LIMIT = 4096      # largest N for which the dense (N,N) state_network is materialised on request


class LazyNetwork:
    """Array-like stand-in for the dense (N,N) state_network: materialised from the engine's CSR only
    when something actually needs the numbers (np.asarray / indexing)."""

    def __init__(self, engine, n_agents, step):
        self._engine, self._n, self._step, self._dense = engine, n_agents, step, None
        self.shape = (n_agents, n_agents)
        self.dtype = np.dtype(np.float64)
        self.ndim = 2

    def _materialise(self):
        if self._dense is None:
            if self._n > DENSE_LIMIT:
                raise MemoryError(f"dense state_network for N={self._n} not materialised (limit {DENSE_LIMIT}); "
                                  "use the engine-backed learner path")
            age = self._engine.step_index - self._step
            if not 0 <= age < self._engine.k:
                raise RuntimeError("state_network of a step that left the engine's K-deep history")
            self._dense = self._engine.network_dense(age=age)[0].astype(np.float64)
        return self._dense

    def __array__(self, dtype=None, copy=None):
        a = self._materialise()
        return a.astype(dtype) if dtype is not None else a

    def __getitem__(self, idx):
        return self._materialise()[idx]

    def reshape(self, *shape):
        return self._materialise().reshape(*shape)

    def diagonal(self, *a, **k):
        return np.zeros(self._n)            # no self loops by construction


class EngineState(tuple):
    """``(state_values, state_network)`` tuple that also remembers which engine / step produced it, so
    the learner-side shims can stay on the device-resident sparse history."""

    def __new__(cls, values, network, engine, step, record_aggregated=False):
        obj = super().__new__(cls, (values, network))
        obj.engine, obj.step = engine, step
        obj.record_aggregated = record_aggregated
        return obj


class FlockingRelativeEnv:
    metadata = {"render.modes": ["human"]}

    def __init__(self):
        self.n_agents = 100
        self.comm_radius = 1.0
        self.comm_radius2 = 1.0
        self.dt = 0.01
        self.v_max = 3.0
        self.v_bias = self.v_max
        self.r_max0 = 1.0
        self.r_max = self.r_max0 * np.sqrt(self.n_agents)
        self.nx_system, self.n_features, self.nu = 4, 6, 2
        self.action_scalar = 10.0
        self.max_accel = 1.0
        self.mean_pooling = True
        self.centralized = True
        self.min_dist_thresh = 0.1
        self.min_degree = 2
        # learner architecture the engine is built for (read from the cfg when present)
        self.k, self.hidden_size, self.n_layers = 3, 32, 2
        self.device_index = 0
        # training loops set this: states then carry the aggregated features the native gradient step consumes
        self.record_aggregated = False
        self._engine = None
        self._engine_key = None
        self._step = 0
        self.x = np.zeros((self.n_agents, self.nx_system))

    # -- configuration ----------------------------------------------------------------------
    def params_from_cfg(self, args):
        self.comm_radius = args.getfloat('comm_radius')
        self.comm_radius2 = self.comm_radius * self.comm_radius
        self.n_agents = args.getint('n_agents')
        self.r_max = self.r_max0 * np.sqrt(self.n_agents)
        self.v_max = args.getfloat('v_max')
        self.v_bias = self.v_max
        self.dt = args.getfloat('dt', fallback=None) or self.dt        # the *_stoch / airsim cfgs carry no dt
        self.k = args.getint('k', fallback=self.k)
        self.hidden_size = args.getint('hidden_size', fallback=self.hidden_size)
        self.n_layers = args.getint('n_layers', fallback=None) or 2
        if args.get('centralized') is not None:
            self.centralized = args.getboolean('centralized')

    def seed(self, seed=None):
        self._seed = seed          # reset() draws from the global numpy RNG (seeded by train.py:27)
        return [seed]

    @property
    def engine(self):
        key = (self.n_agents, self.k, self.hidden_size, self.n_layers, self.comm_radius, self.dt, self.mean_pooling)
        if self._engine is None or key != self._engine_key:
            if self._engine is not None:
                self._engine.close()
            self._engine = FlockEngine(n_agents=self.n_agents, k=self.k, hidden=self.hidden_size,
                                       n_layers=self.n_layers, comm_radius=self.comm_radius, dt=self.dt,
                                       action_scalar=self.action_scalar, mean_pooling=self.mean_pooling,
                                       device=self.device_index, edge_capacity=self._edge_capacity())
            self._engine_key = key
        return self._engine

    def _edge_capacity(self):
        """Directed-edge capacity per agent (mean).  The reference's radius sweeps go up to comm_radius 4 at N = 100
        (cfg/rad.cfg), where nearly every pair is an edge: small flocks get the complete graph, large ones several times
        the expected degree of the reset distribution (a disc of radius N^(1/4): density sqrt(N)/pi) within a memory
        bound; exceeding it raises (``_check_capacity``) instead of returning a void graph."""
        n = self.n_agents
        if n <= 2048:
            return max(n - 1, 1)
        expected = self.comm_radius2 * np.sqrt(n)
        cap = int(min(n - 1, max(64, 3.0 * expected + 32)))
        budget = 8 << 30                                           # bytes for the K-deep CSR ring
        return int(max(16, min(cap, budget // (4 * max(self.k, 1) * n))))

    def _check_capacity(self):
        if self.engine.stats()["overflow"]:
            raise FgnnError(f"edge capacity exceeded ({self._edge_capacity()} per agent at comm_radius {self.comm_radius}): "
                            "the graph of this step is incomplete")

    # -- episode ----------------------------------------------------------------------------
    def _draw_configuration(self, x):
        """One candidate initial configuration into x (N,4), from the global numpy RNG (train.py:27 seeds it)."""
        n = self.n_agents
        length = np.sqrt(np.random.uniform(0, self.r_max, size=(n,)))
        angle = np.pi * np.random.uniform(0, 2, size=(n,))
        x[:, 0] = length * np.cos(angle)
        x[:, 1] = length * np.sin(angle)
        bias = np.random.uniform(low=-self.v_bias, high=self.v_bias, size=(2,))
        x[:, 2] = np.random.uniform(low=-self.v_max, high=self.v_max, size=(n,)) + bias[0]
        x[:, 3] = np.random.uniform(low=-self.v_max, high=self.v_max, size=(n,)) + bias[1]

    def _sample_initial_state(self):
        n = self.n_agents
        x = np.zeros((n, self.nx_system))
        from scipy.spatial import cKDTree
        for _ in range(100000):
            self._draw_configuration(x)
            if n < 2:
                return x
            tree = cKDTree(x[:, 0:2])
            dmin = tree.query(x[:, 0:2], k=2)[0][:, 1].min()
            counts = tree.query_ball_point(x[:, 0:2], self.comm_radius * (1 - 1e-12), return_length=True) - 1
            if counts.min() >= self.min_degree and dmin >= self.min_dist_thresh:
                return x
        raise RuntimeError("no admissible initial configuration found")

    def _observe(self):
        eng = self.engine
        values = eng.get_features().astype(np.float64)
        return EngineState(values, LazyNetwork(eng, self.n_agents, self._step), eng, self._step,
                           self.record_aggregated)

    def _configure_engine(self, engine):
        """Variant hook, called after every engine reset (leader mask, ...)."""
        engine.set_agent_mask(None)

    def _before_step(self, engine):
        """Variant hook, called before every env.step (stochastic time step, ...)."""

    def reset(self, x0=None):
        self.x = self._sample_initial_state() if x0 is None else np.array(x0, dtype=np.float64)
        self._configure_engine(self.engine)
        self.engine.reset(self.x)
        self._check_capacity()
        self._step = 0
        return self._observe()

    def step(self, u):
        u = np.asarray(u)
        assert u.shape == (self.n_agents, self.nu)
        self._before_step(self.engine)
        # float64 actions (env.step(env.controller()), learner/gnn_dagger.py:156-163) are integrated as float64; the
        # policy's ``action.cpu().numpy()`` arrives as fp32 and is widened exactly like numpy does
        reward = self.engine.env_step(u if u.dtype == np.float64 else np.ascontiguousarray(u, dtype=np.float32))
        self._check_capacity()
        self._step += 1
        return self._observe(), float(reward[0]), False, {}

    def controller(self, centralized=None):
        if centralized is None:
            centralized = self.centralized
        return self.engine.controller(centralized=centralized, max_accel=self.max_accel, dtype=np.float64)

    def get_state(self):
        return self.engine.get_state()

    def render(self, mode="human"):
        return None          # plotting is out of scope

    def close(self):
        if self._engine is not None:
            self._engine.close()
            self._engine = None
