"""numpy restatement of the reference learner-side hot path.

TEST INFRASTRUCTURE (see oracle/__init__.py).  Pinned against the reference's own
code + shipped checkpoint through tests/golden/ (oracle/gen_golden.py).

Follows
* learner/state_with_delay.py:22-53  -> ``DelayState``
* learner/actor.py:45-86             -> ``actor_forward``
* learner/gnn_dagger.py:55-72        -> ``select_action``
"""
import numpy as np


class DelayState:
    """State container: fp32 cast, delayed GSO products and delayed feature history.

    delay_gso[0] = I, delay_gso[k] = A_t @ prev.delay_gso[k-1]   (state_with_delay.py:44-47)
    delay_state[0] = x_t, delay_state[k] = prev.delay_state[k-1]  (state_with_delay.py:50-53)
    ``curr_gso`` (state_with_delay.py:38-41) is only built on request: DAGGER never reads it.
    """

    def __init__(self, env_state, prev_state=None, k=3, n_states=6, with_curr_gso=False):
        state_value, state_network = env_state
        n_agents = state_value.shape[0]
        assert state_value.shape == (n_agents, n_states)
        assert state_network.shape == (n_agents, n_agents)
        assert np.sum(np.diag(state_network)) == 0
        self.k = k
        self.values = np.asarray(state_value, dtype=np.float64).T.astype(np.float32).reshape(1, 1, n_states, n_agents)
        self.network = np.asarray(state_network).astype(np.float32).reshape(1, 1, n_agents, n_agents)
        eye = np.eye(n_agents, dtype=np.float32)
        self.curr_gso = None
        if with_curr_gso:
            self.curr_gso = np.zeros((1, k, n_agents, n_agents), dtype=np.float32)
            self.curr_gso[0, 0] = eye
            for i in range(1, k):
                self.curr_gso[0, i] = self.network[0, 0] @ self.curr_gso[0, i - 1]
        self.delay_gso = np.zeros((1, k, n_agents, n_agents), dtype=np.float32)
        self.delay_gso[0, 0] = eye
        if prev_state is not None and k > 1:
            self.delay_gso[0, 1:k] = np.matmul(self.network[0, 0], prev_state.delay_gso[0, 0:k - 1])
        self.delay_state = np.zeros((1, k, n_states, n_agents), dtype=np.float32)
        self.delay_state[0, 0] = self.values[0, 0]
        if prev_state is not None and k > 1:
            self.delay_state[0, 1:k] = prev_state.delay_state[0, 0:k - 1]


def weights_from_state_dict(sd):
    """[(W (out,in,step), b (out,)), ...] from a ``conv_layers.{i}.weight/bias`` mapping
    (layer shapes: learner/actor.py:30-40)."""
    layers = []
    i = 0
    while f"conv_layers.{i}.weight" in sd:
        w = np.asarray(sd[f"conv_layers.{i}.weight"], dtype=np.float32)
        b = np.asarray(sd[f"conv_layers.{i}.bias"], dtype=np.float32)
        layers.append((w.reshape(w.shape[0], w.shape[1], w.shape[2]), b))
        i += 1
    return layers


def aggregate(delay_state, delay_gso):
    """actor.py:68-71: y[b,k,f,n] = sum_m delay_state[b,k,f,m] * delay_gso[b,k,m,n]."""
    return np.matmul(delay_state.astype(np.float32), delay_gso.astype(np.float32))


def actor_forward_any(layers, delay_state, delay_gso, ind_agg):
    """Actor.forward for ANY aggregation index (actor.py:59-86): layers in front of ``ind_agg`` act on every tap k
    separately (kernel (1,1)), the graph aggregation ``x[b,k] @ delay_gso[b,k]`` is applied to the input of layer ``ind_agg``
    (actor.py:68-71), whose kernel (K,1) collapses the taps (actor.py:32-38), the layers behind it are per-agent.

    delay_state (B,K,F,N), delay_gso (B,K,N,N) -> (B,1,n_a,N)."""
    B, K, F, N = delay_state.shape
    assert delay_gso.shape == (B, K, N, N)
    x = np.transpose(delay_state.astype(np.float32), (0, 2, 1, 3))          # (B,F,K,N)   actor.py:63-64
    n_layers = len(layers)
    for i, (w, b) in enumerate(layers):
        if i == ind_agg:                                                    # actor.py:68-71
            x = np.transpose(np.matmul(np.transpose(x, (0, 2, 1, 3)), delay_gso.astype(np.float32)), (0, 2, 1, 3))
        step = w.shape[2]
        assert step == (K if i == ind_agg else 1) and x.shape[2] % step == 0
        if step == 1:
            x = np.einsum('gc,bckn->bgkn', w[:, :, 0], x, dtype=np.float32)
        else:                                                               # kernel (K,1), stride (K,1): one output row
            assert x.shape[2] == K
            x = np.einsum('gck,bckn->bgn', w, x, dtype=np.float32)[:, :, None, :]
        x = (x + b.reshape(1, -1, 1, 1)).astype(np.float32)
        if i < n_layers - 1:                                                # actor.py:75-77
            x = np.tanh(x)
    return x.reshape(B, 1, layers[-1][0].shape[0], N).astype(np.float32)     # actor.py:82


def actor_forward(layers, delay_state, delay_gso, ind_agg=0, return_intermediates=False):
    """Actor.forward for ind_agg == 0 (the only DAGGER setting, gnn_dagger.py:43); other indices: actor_forward_any.

    delay_state (B,K,F,N), delay_gso (B,K,N,N) -> (B,1,n_a,N)."""
    if ind_agg != 0:
        assert not return_intermediates
        return actor_forward_any(layers, delay_state, delay_gso, ind_agg)
    B, K, F, N = delay_state.shape
    assert delay_gso.shape == (B, K, N, N)
    z = aggregate(delay_state, delay_gso)                       # (B,K,F,N)
    w0, b0 = layers[0]
    assert w0.shape[1] == F and w0.shape[2] == K
    x = np.einsum('gfk,bkfn->bgn', w0, z, dtype=np.float32) + b0.reshape(1, -1, 1)
    inter = [z]
    n_layers = len(layers)
    if n_layers > 1:
        x = np.tanh(x.astype(np.float32))
    inter.append(x)
    for i in range(1, n_layers):
        w, b = layers[i]
        x = np.einsum('gh,bhn->bgn', w[:, :, 0], x, dtype=np.float32) + b.reshape(1, -1, 1)
        if i < n_layers - 1:
            x = np.tanh(x.astype(np.float32))
        inter.append(x)
    out = x.reshape(B, 1, layers[-1][0].shape[0], N).astype(np.float32)
    return (out, inter) if return_intermediates else out


def select_action(layers, state):
    """gnn_dagger.py:63-70: (N, n_a) action of the learner for one state container."""
    mu = actor_forward(layers, state.delay_state, state.delay_gso)
    return mu[0, 0].T.copy()
