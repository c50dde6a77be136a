"""Edge-list (sparse) form of the oracle arithmetic, for N too large for dense N x N arrays.

TEST INFRASTRUCTURE (see oracle/__init__.py).  Same arithmetic as oracle.flock_env /
oracle.learner, validated against them at small N in tests/test_oracle.py; the reference
itself has no sparse path.

Candidate pairs come from scipy's cKDTree with a slightly inflated radius; the accept test
is the oracle's own ``dx*dx + dy*dy < R^2`` in float64, so the edge set is identical to the
dense oracle's.
"""
import numpy as np
from scipy.spatial import cKDTree
from scipy import sparse as sp


def radius_edges(x, comm_radius):
    """Directed edge list (i, j), i != j, with dx*dx+dy*dy < R^2 evaluated as the dense oracle does.
    Sorted by (i, j)."""
    pos = np.ascontiguousarray(x[:, 0:2], dtype=np.float64)
    tree = cKDTree(pos)
    pairs = tree.query_pairs(comm_radius * (1.0 + 1e-9) + 1e-12, output_type='ndarray')
    if pairs.size == 0:
        return np.zeros(0, np.int64), np.zeros(0, np.int64)
    i = np.concatenate((pairs[:, 0], pairs[:, 1]))
    j = np.concatenate((pairs[:, 1], pairs[:, 0]))
    dx = x[i, 0] - x[j, 0]
    dy = x[i, 1] - x[j, 1]
    r2 = np.multiply(dx, dx) + np.multiply(dy, dy)
    keep = r2 < comm_radius * comm_radius
    i, j = i[keep], j[keep]
    order = np.lexsort((j, i))
    return i[order], j[order]


def compute_helpers_sparse(x, comm_radius):
    """(state_values (N,6) f64, deg (N,) int64, i, j): features + degrees on the edge list."""
    n = x.shape[0]
    i, j = radius_edges(x, comm_radius)
    d = x[i] - x[j]
    r2 = np.multiply(d[:, 0], d[:, 0]) + np.multiply(d[:, 1], d[:, 1])
    r4 = np.multiply(r2, r2)
    terms = np.stack((d[:, 2], d[:, 0] / r4, d[:, 0] / r2, d[:, 3], d[:, 1] / r4, d[:, 1] / r2), axis=1)
    sv = np.zeros((n, 6))
    for f in range(6):
        sv[:, f] = np.bincount(i, weights=terms[:, f], minlength=n)
    deg = np.bincount(i, minlength=n).astype(np.int64)
    return sv, deg, i, j


def controller_sparse(x, comm_radius, max_accel=1.0, action_scalar=10.0):
    """Decentralised expert controller (flock_env.controller with centralized=False) on the edge list."""
    n = x.shape[0]
    i, j = radius_edges(x, comm_radius)
    d = x[i] - x[j]
    r2 = np.multiply(d[:, 0], d[:, 0]) + np.multiply(d[:, 1], d[:, 1])
    gx = -2.0 * d[:, 0] / (r2 * r2) + 2.0 * d[:, 0] / r2
    gy = -2.0 * d[:, 1] / (r2 * r2) + 2.0 * d[:, 1] / r2
    far = r2 > comm_radius
    gx[far] = 0.0
    gy[far] = 0.0
    ux = -np.bincount(i, weights=gx + d[:, 2], minlength=n)
    uy = -np.bincount(i, weights=gy + d[:, 3], minlength=n)
    lim = max_accel * action_scalar
    return np.clip(np.stack((ux, uy), axis=1), -lim, lim) / action_scalar


def network_csr(n, deg, i, j, mean_pooling=True):
    """state_network as an fp32 scipy CSR: A[i,j] = 1/max(deg_i,1) (row-normalised) or 1."""
    if mean_pooling:
        w = (1.0 / np.maximum(deg, 1).astype(np.float64))[i]
    else:
        w = np.ones(i.shape[0])
    return sp.csr_matrix((w.astype(np.float32), (i, j)), shape=(n, n))


class SparseDelayState:
    """Keeps the last K feature arrays and networks instead of dense GSO products.

    z_0 = x_t;  z_k = x_{t-k} A_t A_{t-1} ... A_{t-k+1}  (newest network applied first;
    learner/state_with_delay.py:44-53 + learner/actor.py:70)."""

    def __init__(self, values_nf, network_csr_f32, prev_state=None, k=3):
        self.k = k
        x = np.asarray(values_nf, dtype=np.float64).astype(np.float32)
        self.hist_x = [x] + ([] if prev_state is None else prev_state.hist_x[:k - 1])
        self.hist_a = [network_csr_f32] + ([] if prev_state is None else prev_state.hist_a[:k - 1])

    def aggregate(self, dtype=np.float32):
        """(K, N, F) array of z_k rows.  dtype=float64 evaluates the same fp32 inputs without rounding
        (a conditioning yardstick for tests, not the reference arithmetic)."""
        n, f = self.hist_x[0].shape
        z = np.zeros((self.k, n, f), dtype=dtype)
        z[0] = self.hist_x[0]
        for k in range(1, self.k):
            if k >= len(self.hist_x):
                break
            y = self.hist_x[k].astype(dtype)
            for a in self.hist_a[:k]:            # A_t first, then A_{t-1}, ...
                y = (a.T.astype(dtype) @ y).astype(dtype)  # (y A)[n] = sum_m A[m,n] y[m]
            z[k] = y
        return z


def readout(layers, z_knf, dtype=np.float32):
    """Per-agent MLP on (K,N,F) aggregated features -> (N, n_a)  (actor.py:73-82).
    dtype=float64: same weights/inputs evaluated without fp32 rounding (conditioning yardstick)."""
    K, N, F = z_knf.shape
    w0, b0 = layers[0]
    x = np.einsum('gfk,knf->ng', w0.astype(dtype), z_knf.astype(dtype), dtype=dtype) + b0.astype(dtype)
    n_layers = len(layers)
    if n_layers > 1:
        x = np.tanh(x.astype(dtype))
    for i in range(1, n_layers):
        w, b = layers[i]
        x = x @ w[:, :, 0].T.astype(dtype) + b.astype(dtype)
        if i < n_layers - 1:
            x = np.tanh(x.astype(dtype))
    return x.astype(dtype)
