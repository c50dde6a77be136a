"""Multi-GPU rollout: agents sharded by index, one process per GPU, one halo all-gather per step.

Rank r owns the contiguous index range [lo_r, lo_r + count_r) and is the only one that integrates those
agents.  Agents interact within comm_radius per hop and a step chains K dependent neighbour gathers, so a
rank can compute its owned agents' actions from the states of every agent within ``K * R`` of its own
region: it recomputes graph, features and hops REDUNDANTLY for those ghosts instead of exchanging
intermediate results.  Per step there is exactly one collective, an all-gather of fixed-capacity
buffers holding the (id, px, py, vx, vy) records of each rank's owned agents that lie inside another
rank's x-window; the header record carries the sender's own x-interval, which is next step's window.
When the index order is spatially sorted (cell-major) the owned range is a strip and the exchange is a
thin boundary layer; for an arbitrary order the windows overlap completely and every agent is sent
(correct, but the slow path -- SURVEY.md section 8e).

``ShardedFlock`` holds the protocol; the per-rank compute sits behind a small backend interface so that
the same orchestration runs on the CUDA engine (``CudaShardBackend``) and, in the CPU tests, on a
numpy backend over gloo.
"""
import numpy as np

RECORD = 5          # doubles per record: id, px, py, vx, vy (record 0 = header: count, x_lo, x_hi, 0, 0)
FAR = 1.0e30        # x coordinate given to agents a rank knows nothing about at reset


def shard_ranges(n_total, world):
    """Contiguous, balanced index ranges: [(lo, count)] * world."""
    base, rem = divmod(n_total, world)
    out, lo = [], 0
    for r in range(world):
        c = base + (1 if r < rem else 0)
        out.append((lo, c))
        lo += c
    return out


def halo_depth(k, comm_radius, margin=0.5):
    """States are needed within K*R of the owned region (K chained neighbour gathers); ``margin`` (in
    units of R) covers motion while an agent's K-deep history becomes valid and the one-step-old windows."""
    return (max(k, 1) + margin) * comm_radius


class CudaShardBackend:
    """Per-rank compute on libfgnn.so (one FlockEngine holding full-size arrays, owning [lo, lo+count))."""

    def __init__(self, n_total, lo, count, ghost_capacity, device=0, **engine_kw):
        import torch
        from multiagent_gnn_policies_b200.engine import FlockEngine
        self.torch = torch
        self.engine = FlockEngine(n_agents=n_total, device=device, shard_lo=lo, shard_count=count,
                                  ghost_capacity=ghost_capacity, **engine_kw)
        self.device = self.engine.device
        self.n_total, self.lo, self.count = n_total, lo, count

    def new_buffer(self, rows):
        return self.torch.zeros((rows, RECORD), dtype=self.torch.float64, device=self.device)

    def reset(self, x_global):
        self.engine.reset(x_global)

    def local_step(self):
        self.engine.shard_local_step()

    def policy(self, out):
        """select_action for the owned agents into ``out`` (owned slice; host or device)."""
        return self.engine.policy(out=out)

    def integrate(self, u):
        """first half of env.step for the owned agents from ``u`` (owned slice; host or device)."""
        self.engine.integrate(u)

    def pack(self, windows, window_stride, world, rank, depth, send, cap):
        self.engine.shard_pack(windows, window_stride, world, rank, depth, send, cap)

    def unpack(self, recv, world, rank, cap, depth):
        self.engine.shard_unpack(recv, world, rank, cap, depth)

    def build(self, advance):
        self.engine.build_graph(advance=advance)

    # CUDA-graph replayed halves of a step (same kernels as local_step+pack / unpack+build)
    def step_begin(self, windows, window_stride, world, rank, depth, send, cap):
        self.engine.shard_step_begin(windows, window_stride, world, rank, depth, send, cap)

    def step_end(self, recv, world, rank, cap, depth):
        self.engine.shard_step_end(recv, world, rank, cap, depth)

    def owned_state(self):
        return self.engine.get_state()[self.lo:self.lo + self.count]

    def owned_action(self):
        return self.engine.get_action()[self.lo:self.lo + self.count]

    def overflow(self):
        return self.engine.stats()["overflow"]


class ShardedFlock:
    """The halo protocol of one rank.  ``all_gather(send) -> recv`` concatenates every rank's buffer
    (torch.distributed over NCCL in production, gloo or an in-process list in tests)."""

    def __init__(self, backend, rank, world, k, comm_radius, capacity, all_gather, margin=0.5, send_slack=0.25):
        self.backend, self.rank, self.world = backend, rank, world
        self.cap = int(capacity)
        self.depth = halo_depth(k, comm_radius, margin)
        self.send_depth = self.depth + send_slack * comm_radius      # windows are one step old when used
        self.all_gather = all_gather
        self.send = backend.new_buffer(self.cap + 1)
        self.recv = None
        self.windows0 = backend.new_buffer(world)                    # [world][RECORD], cols 1,2 = lo, hi

    def reset(self, x_global, ranges):
        """``x_global`` (n_total,4) f64: the rank must know the true state of every agent within the halo
        depth of its own strip; agents it knows nothing about must sit at x = FAR."""
        x_global = np.ascontiguousarray(x_global, dtype=np.float64)
        self.backend.reset(x_global)
        win = np.zeros((self.world, RECORD))
        for q, (lo, cnt) in enumerate(ranges):
            xs = x_global[lo:lo + cnt, 0]
            xs = xs[xs < 0.5 * FAR]
            # a rank that cannot see another rank's agents uses an empty window for it
            win[q, 1], win[q, 2] = (xs.min(), xs.max()) if xs.size else (FAR, -FAR)
        self._upload(self.windows0, win)
        self._exchange(self.windows0, RECORD)
        self.backend.build(False)

    def _upload(self, dst, arr):
        if hasattr(dst, "copy_"):
            import torch
            dst.copy_(torch.from_numpy(arr))
        else:
            dst[...] = arr

    def _exchange(self, windows, stride):
        # windows[q*stride + 1 .. 2] = rank q's x-interval: pass a view that starts at column 1
        win_view = windows.reshape(-1)[1:]
        self.backend.pack(win_view, stride, self.world, self.rank, self.send_depth, self.send, self.cap)
        self.recv = self.all_gather(self.send)
        self.backend.unpack(self.recv, self.world, self.rank, self.cap, self.depth)

    def step(self):
        """One closed-loop step: local policy + integrator for owned agents, halo exchange, rebuild."""
        if hasattr(self.backend, "step_begin"):          # CUDA backend: two graph launches around the all-gather
            win_view = self.recv.reshape(-1)[1:]
            self.backend.step_begin(win_view, (self.cap + 1) * RECORD, self.world, self.rank, self.send_depth,
                                    self.send, self.cap)
            self.recv = self.all_gather(self.send)
            self.backend.step_end(self.recv, self.world, self.rank, self.cap, self.depth)
            return
        self.backend.local_step()
        self._exchange(self.recv, (self.cap + 1) * RECORD)
        self.backend.build(True)

    def step_host(self, action_host):
        """The same step through host buffers, as the reference loop does it (learner/gnn_dagger.py:196-201):
        select_action -> host array -> env.step(host array).  ``action_host``: (count, 2) fp32, pinned."""
        self.backend.policy(action_host)          # D2H (synchronises)
        self.backend.integrate(action_host)       # H2D
        self._exchange(self.recv, (self.cap + 1) * RECORD)
        self.backend.build(True)


def nccl_all_gather(world, cap, device):
    """all_gather closure on torch.distributed (default group); buffers are float64 CUDA tensors."""
    import torch
    import torch.distributed as dist
    recv = torch.zeros((world, cap + 1, RECORD), dtype=torch.float64, device=device)

    def gather(send):
        dist.all_gather_into_tensor(recv.view(-1), send.view(-1))
        return recv
    return gather
