"""numpy/oracle implementation of the per-rank backend interface of parallel.ShardedFlock.
TEST INFRASTRUCTURE: lets the halo + hand-over protocol (what is sent, windows, redundant ghost compute,
halo depth, ownership transfer) run on CPU over gloo and be compared with the single-process oracle."""
import numpy as np
from scipy import sparse as sp

from oracle import flock_env, sparse
from multiagent_gnn_policies_b200.parallel import RECORD


class NumpyShardBackend:
    def __init__(self, n_total, lo, count, layers, k=3, comm_radius=1.0, dt=0.01):
        self.n, self.lo, self.count = n_total, lo, count
        self.layers, self.k, self.R, self.dt = layers, k, comm_radius, dt
        self.x = np.zeros((n_total, 4))
        self.own = np.arange(lo, lo + count)
        self.ghosts = np.zeros(0, np.int64)
        self.hist_x, self.hist_a = [], []
        self.action = np.zeros((n_total, 2), np.float32)
        self.packed_overflow = False
        self.t = 0
        self.shift = 0.0
        self.handed_over = 0

    def new_buffer(self, rows):
        return np.zeros((rows, RECORD))

    def configure(self, bounds, world, rank, depth, margin, dshift, handover_after):
        self.bounds, self.world, self.rank = np.asarray(bounds), world, rank
        self.depth, self.margin, self.dshift, self.handover_after = depth, margin, dshift, handover_after

    def reset(self, x_global):
        self.x = np.array(x_global, dtype=np.float64)
        self.own = np.arange(self.lo, self.lo + self.count)
        self.ghosts = np.zeros(0, np.int64)
        self.hist_x, self.hist_a = [], []
        self.t, self.shift = 0, 0.0

    @property
    def pool(self):
        return np.concatenate([self.own, self.ghosts])

    def build(self, advance):
        """graph + features over the agents present (pool); rows of absent agents stay empty."""
        if advance:
            self.t += 1
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
        own = self.own
        act = sparse.readout(self.layers, z[:, own])
        self.action[own] = act
        self.x[own] = flock_env.integrate(self.x[own], act, self.dt)

    def _strip(self, xs):
        return (xs[:, None] >= self.bounds[None, 1:self.world]).sum(axis=1)

    def pack(self, windows, stride, send, cap, advance):
        if advance:
            self.shift += self.dshift
        own = self.own
        x = self.x[own, 0]
        xs = x - self.shift
        new_owner = np.full(own.size, -1)
        if self.t >= self.handover_after:
            st = self._strip(xs)
            right = (st > self.rank) & (xs - self.bounds[self.rank + 1] > self.margin)
            left = (st < self.rank) & (self.bounds[self.rank] - xs > self.margin)
            new_owner[right | left] = st[right | left]
        give = new_owner >= 0
        wanted = give.copy()
        for q in range(self.world):
            if q == self.rank:
                continue
            lo, hi = windows[q * stride], windows[q * stride + 1]
            in_strip = (xs >= self.bounds[q] - self.depth) & (xs <= self.bounds[q + 1] + self.depth)
            in_ival = (x >= lo - self.depth) & (x <= hi + self.depth)
            wanted |= in_strip | in_ival
        ids = own[wanted]
        send[...] = 0
        kept = own[~give]
        send[0, 0] = ids.size
        if kept.size:
            send[0, 1], send[0, 2] = self.x[kept, 0].min(), self.x[kept, 0].max()
        else:
            send[0, 1], send[0, 2] = 1e300, -1e300
        n = min(ids.size, cap)
        send[1:n + 1, 0] = ids[:n]
        send[1:n + 1, 1:5] = self.x[ids[:n]]
        send[1:n + 1, 5] = new_owner[wanted][:n]
        self._own_new = kept
        self.ghosts = own[give]
        self.handed_over += int(give.sum())

    def unpack(self, recv, cap):
        recv = np.asarray(recv).reshape(self.world, cap + 1, RECORD)
        lo, hi = recv[self.rank, 0, 1] - self.depth, recv[self.rank, 0, 2] + self.depth
        new_own, ghosts = [self._own_new], [self.ghosts]
        for q in range(self.world):
            if q == self.rank:
                continue
            cnt = int(recv[q, 0, 0])
            if cnt > cap:
                self.packed_overflow = True
            rec = recv[q, 1:min(cnt, cap) + 1]
            xs = rec[:, 1] - self.shift
            in_strip = (xs >= self.bounds[self.rank] - self.depth) & (xs <= self.bounds[self.rank + 1] + self.depth)
            in_ival = (rec[:, 1] >= lo) & (rec[:, 1] <= hi)
            mine = rec[:, 5].astype(np.int64) == self.rank
            rec = rec[mine | in_strip | in_ival]
            mine = rec[:, 5].astype(np.int64) == self.rank
            ids = rec[:, 0].astype(np.int64)
            self.x[ids] = rec[:, 1:5]
            new_own.append(ids[mine])
            ghosts.append(ids[~mine])
        self.own = np.concatenate(new_own)
        self.ghosts = np.concatenate(ghosts)

    def owned(self):
        return self.own.copy()

    def owned_state(self):
        ids = np.sort(self.own)
        return ids, self.x[ids]

    def owned_action(self):
        ids = np.sort(self.own)
        return ids, self.action[ids]

    def overflow(self):
        return self.packed_overflow
