"""Delayed-state container with the reference's signature and public attributes
(learner/state_with_delay.py:6-53): ``values, network, curr_gso, delay_gso, delay_state``.

When ``env_state`` comes from the engine-backed env (an ``EngineState``) nothing dense is built: the
K-deep history already lives on the device as CSR graphs + feature rows, and ``DAGGER.select_action``
runs the sparse kernels.  When the env is in training mode (``env.record_aggregated``, set by
``train_dagger`` / ``train_cloning``) the state also captures ``aggregated``: the K-hop aggregated features
z (K,N,6) on the device -- all that ``gradient_step`` needs from a stored state (ind_agg = 0), 6K floats per
agent instead of the reference's dense (K,N,N) operator.  The dense attributes are materialised lazily, on first access, from short
per-state histories of device tensors (never from ``prev_state`` itself, so episodes do not leak) --
that is what the replay buffer / ``gradient_step`` read.  A plain ``(ndarray, ndarray)`` tuple (e.g.
from another env) takes the dense route immediately, like the reference.
"""
import numpy as np
import torch

DENSE_LIMIT = 4096


class MultiAgentStateWithDelay(object):

    def __init__(self, device, args, env_state, prev_state=None, k=None):
        n_states = args.getint('n_states')
        n_agents = args.getint('n_agents')
        k = k or args.getint('k')
        self.k, self.n_states, self.n_agents = k, n_states, n_agents
        self.device = torch.device(device)
        if self.device.type != "cuda":
            raise RuntimeError("MultiAgentStateWithDelay needs a CUDA device: this build has no CPU fallback")
        state_value, state_network = env_state
        assert state_value.shape == (n_agents, n_states)
        assert state_network.shape == (n_agents, n_agents)

        self.engine = getattr(env_state, "engine", None)
        self.step = getattr(env_state, "step", None)
        self.aggregated = None
        if (self.engine is not None and getattr(env_state, "record_aggregated", False)
                and self.engine.step_index == self.step and self.engine.k == k):
            self.aggregated = self.engine.aggregate()          # (K,N,6) CUDA tensor, hops on the sparse history
        self._values = self._network = self._curr_gso = self._delay_gso = self._delay_state = None

        x_t = torch.as_tensor(np.asarray(state_value), dtype=torch.float32, device=self.device).t().contiguous()   # (F,N)
        if self.engine is not None:
            a_t = None                                   # exported from the engine only if someone asks
            if n_agents <= DENSE_LIMIT:
                a_t = self.engine.network_dense(age=0, device=True)[0]                 # (N,N) device tensor, one kernel
        else:
            net = np.asarray(state_network)
            assert np.sum(np.diag(net)) == 0             # assume no self loops (state_with_delay.py:26)
            a_t = torch.as_tensor(net, dtype=torch.float32, device=self.device)
        # short histories of device tensors: newest first, at most K entries
        prev_x = prev_state._x_hist if prev_state is not None else ()
        prev_a = prev_state._a_hist if prev_state is not None else ()
        self._x_hist = ((x_t,) + prev_x)[:k]
        self._a_hist = ((a_t,) + prev_a)[:max(k - 1, 1)]
        self._age = 0 if prev_state is None else min(prev_state._age + 1, k)    # how many earlier states exist

    # -- lazily materialised dense views ----------------------------------------------------
    def _need_dense(self):
        if any(a is None for a in self._a_hist):
            raise MemoryError(f"dense GSO for N={self.n_agents} is not materialised (limit {DENSE_LIMIT})")

    @property
    def values(self):
        if self._values is None:
            self._values = self._x_hist[0].view(1, 1, self.n_states, self.n_agents)
        return self._values

    @property
    def network(self):
        if self._network is None:
            self._need_dense()
            self._network = self._a_hist[0].view(1, 1, self.n_agents, self.n_agents)
        return self._network

    @property
    def curr_gso(self):
        if self._curr_gso is None:
            self._need_dense()
            n, a = self.n_agents, self._a_hist[0]
            g = torch.zeros((1, self.k, n, n), device=self.device)
            g[0, 0] = torch.eye(n, device=self.device)
            for i in range(1, self.k):
                g[0, i] = torch.matmul(a, g[0, i - 1])
            self._curr_gso = g
        return self._curr_gso

    @property
    def delay_gso(self):
        """[I, A_t, A_t A_{t-1}, ...]; slices beyond the episode start are zero (state_with_delay.py:44-47)."""
        if self._delay_gso is None:
            self._need_dense()
            n = self.n_agents
            g = torch.zeros((1, self.k, n, n), device=self.device)
            g[0, 0] = torch.eye(n, device=self.device)
            for i in range(1, self.k):
                if i > self._age:
                    break
                # delay_gso_t[i] = A_t @ delay_gso_{t-1}[i-1] = A_t (A_{t-1} (... A_{t-i+1})), same association
                prod = self._a_hist[i - 1]
                for j in range(i - 2, -1, -1):
                    prod = torch.matmul(self._a_hist[j], prod)
                g[0, i] = prod
            self._delay_gso = g
        return self._delay_gso

    @property
    def delay_state(self):
        if self._delay_state is None:
            s = torch.zeros((1, self.k, self.n_states, self.n_agents), device=self.device)
            for i, x in enumerate(self._x_hist):
                s[0, i] = x
            self._delay_state = s
        return self._delay_state
