"""The other flocking environments the reference's cfgs name, on the same CUDA engine:

* ``FlockingLeader-v0``     (cfg/dagger_leader.cfg:24, cfg/vel_leader_baseline.cfg)
* ``FlockingTwoFlocks-v0``  (cfg/dagger_twoflocks.cfg:24, cfg/n_twoflocks.cfg; flocking_gym_test.py:6)
* ``FlockingStochastic-v0`` (cfg/dagger_stoch.cfg:24, cfg/rad_stoch.cfg, cfg/transfer_stoch.cfg)

gym_flock is an un-vendored, un-pinned dependency of the reference (README.md:7): what distinguishes the
variants is restated from memory of the upstream package [UNVERIFIED-MEMORY], every constant is an
attribute, and oracle.flock_env holds the same semantics for the parity tests (parity unpinned, like the
base env).  Kernels are shared with FlockingRelative-v0: the leader mask lives in the integrator
(fgnn_set_agent_mask), the random time step is installed per step (fgnn_set_dt), two flocks differ only in
the host-side initial draw.
"""
import numpy as np

from gym_flock.envs.flocking_relative import FlockingRelativeEnv


class FlockingLeaderEnv(FlockingRelativeEnv):
    """The first ``n_leaders`` agents share one constant velocity and ignore every action (``u * mask``)."""

    def __init__(self):
        super().__init__()
        self.n_leaders = 2
        self.mask = np.ones((self.n_agents,))
        self.mask[0:self.n_leaders] = 0

    def params_from_cfg(self, args):
        super().params_from_cfg(args)
        self.mask = np.ones((self.n_agents,))
        self.mask[0:self.n_leaders] = 0

    def _configure_engine(self, engine):
        engine.set_agent_mask(self.mask)

    def _sample_initial_state(self):
        x = super()._sample_initial_state()
        x[0:self.n_leaders, 2:4] = np.ones((self.n_leaders, 2)) * np.random.uniform(
            low=-self.v_max, high=self.v_max, size=(1, 1))
        return x


class FlockingTwoFlocksEnv(FlockingRelativeEnv):
    """Two half-flocks (discs of half the area: same density), centres ``flock_offset`` apart along x
    (default: tangent discs), opposite velocity biases."""

    def __init__(self):
        super().__init__()
        self.flock_offset = None

    def _draw_configuration(self, x):
        n = self.n_agents
        half = n // 2
        offset = 2.0 * np.sqrt(0.5 * self.r_max) if self.flock_offset is None else self.flock_offset
        length = np.sqrt(np.random.uniform(0, 0.5 * self.r_max, size=(n,)))
        angle = np.pi * np.random.uniform(0, 2, size=(n,))
        x[:, 0] = length * np.cos(angle)
        x[:, 1] = length * np.sin(angle)
        x[:half, 0] -= 0.5 * offset
        x[half:, 0] += 0.5 * offset
        bias = np.random.uniform(low=-self.v_bias, high=self.v_bias, size=(2,))
        sign = np.where(np.arange(n) < half, 1.0, -1.0)
        x[:, 2] = np.random.uniform(low=-self.v_max, high=self.v_max, size=(n,)) + sign * bias[0]
        x[:, 3] = np.random.uniform(low=-self.v_max, high=self.v_max, size=(n,)) + sign * bias[1]


class FlockingStochasticEnv(FlockingRelativeEnv):
    """Random time step per env.step: dt ~ max(N(dt_mean, dt_sigma), dt_min) (the *_stoch cfgs carry no ``dt``)."""

    def __init__(self):
        super().__init__()
        self.dt_mean, self.dt_sigma, self.dt_min = 0.1, 0.02, 1e-3
        self.dt = self.dt_mean

    def params_from_cfg(self, args):
        super().params_from_cfg(args)
        if args.get('dt') is not None:
            self.dt_mean = args.getfloat('dt')
        self.dt = self.dt_mean

    def draw_dt(self):
        return float(max(np.random.normal(self.dt_mean, self.dt_sigma), self.dt_min))

    @property
    def engine(self):
        # the engine is keyed on dt in the base class; here dt changes every step, so key on the mean
        dt, self.dt = self.dt, self.dt_mean
        try:
            return FlockingRelativeEnv.engine.fget(self)
        finally:
            self.dt = dt

    def _before_step(self, engine):
        self.dt = self.draw_dt()
        engine.set_dt(self.dt)
