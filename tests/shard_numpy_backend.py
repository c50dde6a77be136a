"""numpy/oracle implementation of the per-rank backend interface of parallel.ShardedFlock.
TEST INFRASTRUCTURE: lets the halo protocol (what is sent, window logic, redundant ghost compute, halo
depth) run on CPU over gloo and be compared with the single-process oracle."""
import numpy as np
from scipy import sparse as sp

from oracle import flock_env, sparse
from multiagent_gnn_policies_b200.parallel import RECORD


class NumpyShardBackend:
    def __init__(self, n_total, lo, count, layers, k=3, comm_radius=1.0, dt=0.01):
        self.n, self.lo, self.count = n_total, lo, count
        self.layers, self.k, self.R, self.dt = layers, k, comm_radius, dt
        self.x = np.zeros((n_total, 4))
        self.pool = np.arange(lo, lo + count)
        self.hist_x, self.hist_a = [], []
        self.action = np.zeros((n_total, 2), np.float32)
        self.packed_overflow = False

    def new_buffer(self, rows):
        return np.zeros((rows, RECORD))

    def reset(self, x_global):
        self.x = np.array(x_global, dtype=np.float64)
        self.pool = np.arange(self.lo, self.lo + self.count)
        self.hist_x, self.hist_a = [], []

    def build(self, advance):
        """graph + features over the agents present (pool); rows of absent agents stay empty."""
        pool = np.sort(self.pool)
        sv, deg, i, j = sparse.compute_helpers_sparse(self.x[pool], self.R)
        feats = np.zeros((self.n, 6), np.float32)
        feats[pool] = sv.astype(np.float32)
        w = (1.0 / np.maximum(deg, 1).astype(np.float64))[i].astype(np.float32)
        a = sp.csr_matrix((w, (pool[i], pool[j])), shape=(self.n, self.n))
        self.hist_x = ([feats] + self.hist_x)[:self.k]
        self.hist_a = ([a] + self.hist_a)[:self.k]

    def local_step(self):
        z = np.zeros((self.k, self.n, 6), np.float32)
        z[0] = self.hist_x[0]
        for k in range(1, self.k):
            if k >= len(self.hist_x):
                break
            y = self.hist_x[k]
            for a in self.hist_a[:k]:
                y = (a.T @ y).astype(np.float32)
            z[k] = y
        own = slice(self.lo, self.lo + self.count)
        act = sparse.readout(self.layers, z[:, own])
        self.action[own] = act
        self.x[own] = flock_env.integrate(self.x[own], act, self.dt)

    def pack(self, windows, stride, world, rank, depth, send, cap):
        own = np.arange(self.lo, self.lo + self.count)
        xs = self.x[own, 0]
        wanted = np.zeros(own.size, bool)
        for q in range(world):
            if q == rank:
                continue
            lo, hi = windows[q * stride], windows[q * stride + 1]
            wanted |= (xs >= lo - depth) & (xs <= hi + depth)
        ids = own[wanted]
        send[...] = 0
        send[0, :3] = (ids.size, xs.min(), xs.max())
        n = min(ids.size, cap)
        send[1:n + 1, 0] = ids[:n]
        send[1:n + 1, 1:] = self.x[ids[:n]]
        self.pool = own

    def unpack(self, recv, world, rank, cap, depth):
        recv = np.asarray(recv).reshape(world, cap + 1, RECORD)
        lo, hi = recv[rank, 0, 1] - depth, recv[rank, 0, 2] + depth
        ghosts = []
        for q in range(world):
            if q == rank:
                continue
            cnt = int(recv[q, 0, 0])
            if cnt > cap:
                self.packed_overflow = True
            rec = recv[q, 1:min(cnt, cap) + 1]
            rec = rec[(rec[:, 1] >= lo) & (rec[:, 1] <= hi)]
            ids = rec[:, 0].astype(np.int64)
            self.x[ids] = rec[:, 1:]
            ghosts.append(ids)
        self.pool = np.concatenate([np.arange(self.lo, self.lo + self.count)] + ghosts)

    def owned_state(self):
        return self.x[self.lo:self.lo + self.count]

    def owned_action(self):
        return self.action[self.lo:self.lo + self.count]

    def overflow(self):
        return self.packed_overflow
