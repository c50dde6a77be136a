"""Multi-GPU rollout: one process per GPU, one halo all-gather per step, ownership hand-over.

Every rank keeps full-size arrays (global agent indices).  Rank r starts owning a contiguous index range and
is the only rank that integrates what it owns.  A step chains K dependent neighbour gathers, so a rank can
compute its owned agents' actions from the states of every agent within ``K * R`` of them: it recomputes
graph, features and hops REDUNDANTLY for those ghosts instead of exchanging intermediate results.  Per step
there is exactly one collective, an all-gather of fixed-capacity buffers with the records
``(id, px, py, vx, vy, new_owner)`` of each rank's owned agents that lie inside another rank's window; the
header record carries the sender's own x-interval (part of next step's windows).

Territories are x-strips in a frame moving with the flock.  When an owned agent has entered another rank's
strip by more than ``margin`` its owner hands it over: the receiver already holds it as a ghost with a valid
K-deep history, so the hand-over moves no data -- the record just names the new owner.  Owned sets stay
spatially compact however long the rollout runs, and an arbitrary initial index order re-partitions itself.

``ShardedFlock`` holds the protocol; the per-rank compute sits behind a small backend interface so the same
orchestration runs on the CUDA engine (``CudaShardBackend``) and, in the CPU tests, on a numpy backend
over gloo.
"""
import numpy as np

RECORD = 6          # doubles per record: id, px, py, vx, vy, new_owner (record 0 = header: count, x_lo, x_hi, 0, 0, 0)
FAR = 1.0e30        # x coordinate given to agents a rank knows nothing about at reset
INF = 1.0e300


def bind_to_local_cpus(device_index):
    """One process per GPU: run this rank on the CPUs NVML names as local to its GPU, so that its pinned host buffers
    (first touch) and the driver's copies stay on the GPU's own NUMA node.  With eight ranks moving 8 MB down and 8 MB up per
    step through host buffers, a rank floating on the other socket sends all of it across the inter-socket link.  Best
    effort: returns the sorted CPU list it bound to, or None when NVML is missing, the mask is empty inside this process's
    cpuset, or it already equals the allowed set (then nothing is changed)."""
    import os
    try:
        import pynvml
        import torch
        pynvml.nvmlInit()
        uuid = "GPU-" + str(torch.cuda.get_device_properties(device_index).uuid)
        try:
            handle = pynvml.nvmlDeviceGetHandleByUUID(uuid)
        except Exception:
            handle = pynvml.nvmlDeviceGetHandleByUUID(uuid.encode())
        allowed = os.sched_getaffinity(0)
        words = pynvml.nvmlDeviceGetCpuAffinity(handle, (max(allowed) + 64) // 64)
        local = {64 * w + b for w, word in enumerate(words) for b in range(64) if (int(word) >> b) & 1} & allowed
        if not local or local == allowed:
            return None
        os.sched_setaffinity(0, local)
        return sorted(local)
    except Exception:
        return None


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
    """States are needed within K*R of the owned agents (K chained neighbour gathers); ``margin`` (in units
    of R) covers motion while an agent's K-deep history becomes valid and the one-step-old windows."""
    return (max(k, 1) + margin) * comm_radius


def strip_bounds(x_global, ranges):
    """Territory boundaries from a fully known initial state: midpoints between consecutive ranks' x-extents
    (only meaningful when the index order is spatially sorted along x).  (world + 1,) with +-INF at the ends."""
    world = len(ranges)
    b = np.empty(world + 1)
    b[0], b[world] = -INF, INF
    for q in range(1, world):
        lo_prev, c_prev = ranges[q - 1]
        lo, c = ranges[q]
        b[q] = 0.5 * (x_global[lo_prev:lo_prev + c_prev, 0].max() + x_global[lo:lo + c, 0].min())
    if np.any(np.diff(b[1:world]) <= 0):          # not sorted along x: equal-width strips over the extent instead
        xs = x_global[:, 0]
        b[1:world] = np.linspace(xs.min(), xs.max(), world + 1)[1:world]
    return b


class CudaShardBackend:
    """Per-rank compute on libfgnn.so (one FlockEngine holding full-size arrays, initially owning [lo, lo+count))."""

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

    def configure(self, bounds, world, rank, depth, margin, dshift, handover_after):
        self.engine.shard_configure(bounds, world, rank, depth, margin, dshift, handover_after)

    def reset(self, x_global):
        self.engine.reset(x_global)

    def local_step(self):
        self.engine.shard_local_step()

    def policy(self, out):
        """select_action for the owned agents into ``out`` (owned-list order, list-capacity rows; host or device)."""
        return self.engine.policy(out=out)

    def integrate(self, u):
        """first half of env.step for the owned agents from ``u`` (same layout as ``policy``)."""
        self.engine.integrate(u)

    def pack(self, windows, window_stride, send, cap, advance):
        self.engine.shard_pack(windows, window_stride, send, cap, advance)

    def unpack(self, recv, cap):
        self.engine.shard_unpack(recv, cap)

    def build(self, advance):
        self.engine.build_graph(advance=advance)

    # CUDA-graph replayed halves of a step (same kernels as local_step+pack / unpack+build)
    def step_begin(self, windows, window_stride, send, cap):
        self.engine.shard_step_begin(windows, window_stride, send, cap)

    def step_end(self, recv, cap):
        self.engine.shard_step_end(recv, cap)

    def init_comm(self, rank, world):
        """Give the engine its own NCCL communicator (id broadcast through torch.distributed): the step then
        runs as ONE CUDA graph with the all-gather inside (``step_fused``)."""
        import torch.distributed as dist
        from multiagent_gnn_policies_b200.engine import comm_unique_id
        box = [comm_unique_id() if rank == 0 else None]
        dist.broadcast_object_list(box, src=0)
        self.engine.comm_init(box[0], rank, world)
        self.native_comm = True

    def step_fused(self, send, recv, cap):
        self.engine.shard_step(send, recv, cap)

    # p2p transport: records go straight into the peers' inboxes (CUDA-IPC mappings over NVLink); no collective in the step
    def p2p_alloc(self, world, rank, cap):
        return self.engine.p2p_alloc(world, rank, cap)

    def p2p_connect(self, handles=None, direct_ptrs=None):
        self.engine.p2p_connect(handles=handles, direct_ptrs=direct_ptrs)
        self.p2p = True

    def p2p_seed(self, gathered):
        self.engine.p2p_seed(gathered)

    def step_p2p(self):
        self.engine.shard_step_p2p()

    def exchange_p2p(self, advance):
        self.engine.shard_exchange_p2p(advance)

    def owned(self):
        return self.engine.shard_owned()

    def owned_state(self):
        ids = np.sort(self.owned())
        return ids, self.engine.get_state()[ids]

    def owned_action(self):
        ids = np.sort(self.owned())
        return ids, self.engine.get_action()[ids]

    def overflow(self):
        return self.engine.stats()["overflow"]


class ShardedFlock:
    """The halo / hand-over protocol of one rank.  ``all_gather(send) -> recv`` concatenates every rank's
    buffer (torch.distributed over NCCL in production, gloo or an in-process stack in tests)."""

    def __init__(self, backend, rank, world, k, comm_radius, capacity, all_gather, dt=0.01, margin=0.5,
                 send_slack=0.25, handover_margin=1.0):
        self.backend, self.rank, self.world, self.k, self.dt = backend, rank, world, k, dt
        self.cap = int(capacity)
        self.R = comm_radius
        self.depth = halo_depth(k, comm_radius, margin) + send_slack * comm_radius   # windows are one step old
        self.handover_margin = handover_margin * comm_radius
        self.all_gather = all_gather
        self.send = backend.new_buffer(self.cap + 1)
        self.recv = None
        self.windows0 = backend.new_buffer(world)                    # [world][RECORD], cols 1,2 = lo, hi

    def reset(self, x_global, ranges, bounds=None, frame_velocity=0.0):
        """``x_global`` (n_total,4) f64: the rank must know the true state of every agent within the halo
        depth of its own agents; agents it knows nothing about must sit at x = FAR.  ``bounds``: territory
        boundaries (world+1,), identical on every rank; default: from ``x_global`` (needs full knowledge).
        ``frame_velocity``: x-velocity of the frame the territories move with (the flock's mean vx)."""
        x_global = np.ascontiguousarray(x_global, dtype=np.float64)
        if bounds is None:
            bounds = strip_bounds(x_global, ranges)
        self.bounds = np.asarray(bounds, dtype=np.float64)
        self.backend.configure(self.bounds, self.world, self.rank, self.depth, self.handover_margin,
                               frame_velocity * self.dt, self.k + 1)
        self.backend.reset(x_global)
        win = np.zeros((self.world, RECORD))
        for q, (lo, cnt) in enumerate(ranges):
            xs = x_global[lo:lo + cnt, 0]
            xs = xs[xs < 0.5 * FAR]
            # a rank that cannot see another rank's agents uses an empty window for it
            win[q, 1], win[q, 2] = (xs.min(), xs.max()) if xs.size else (FAR, -FAR)
        self._upload(self.windows0, win)
        self._exchange(self.windows0, RECORD, advance=False)
        self.backend.build(False)

    def _upload(self, dst, arr):
        if hasattr(dst, "copy_"):
            import torch
            dst.copy_(torch.from_numpy(arr))
        else:
            dst[...] = arr

    def _exchange(self, windows, stride, advance):
        # windows[q*stride + 1 .. 2] = rank q's x-interval: pass a view that starts at column 1
        win_view = windows.reshape(-1)[1:]
        self.backend.pack(win_view, stride, self.send, self.cap, advance)
        self.recv = self.all_gather(self.send)
        self.backend.unpack(self.recv, self.cap)

    def enable_p2p(self, all_gather_object):
        """Switch the halo of the following steps to peer-to-peer stores (call once, right after ``reset``):
        every rank allocates its inbox, the CUDA-IPC handles are exchanged with ``all_gather_object(obj) -> list`` (e.g.
        torch.distributed.all_gather_object), the peers' inboxes are mapped, and the buffer gathered by the reset-time
        exchange seeds the inbox headers.  The ranks must be processes of one node whose GPUs have peer access."""
        handle, _ = self.backend.p2p_alloc(self.world, self.rank, self.cap)
        handles = all_gather_object(handle)
        self.backend.p2p_connect(handles=b"".join(handles))
        self.backend.p2p_seed(self.recv)
        self.backend.engine.sync()
        all_gather_object(b"ready")               # every inbox is mapped and seeded before anybody's first step stores into it

    def step(self):
        """One closed-loop step: local policy + integrator for owned agents, halo exchange, rebuild."""
        stride = (self.cap + 1) * RECORD
        if getattr(self.backend, "p2p", False):          # one CUDA graph, halo records stored straight into the peers' inboxes
            self.backend.step_p2p()
            return
        if getattr(self.backend, "native_comm", False):  # CUDA backend with its own communicator: one graph per step
            self.backend.step_fused(self.send, self.recv, self.cap)
            return
        if hasattr(self.backend, "step_begin"):          # CUDA backend: two graph launches around the all-gather
            self.backend.step_begin(self.recv.reshape(-1)[1:], stride, self.send, self.cap)
            self.recv = self.all_gather(self.send)
            self.backend.step_end(self.recv, self.cap)
            return
        self.backend.local_step()
        self._exchange(self.recv, stride, advance=True)
        self.backend.build(True)

    def step_host(self, action_host):
        """The same step through host buffers, as the reference loop does it (learner/gnn_dagger.py:196-201):
        select_action -> host array -> env.step(host array).  ``action_host``: (list capacity, 2) fp32, pinned."""
        self.backend.policy(action_host)          # D2H (synchronises)
        self.backend.integrate(action_host)       # H2D
        if getattr(self.backend, "p2p", False):   # halo records stored straight into the peers' inboxes: no collective
            self.backend.exchange_p2p(True)
        else:
            self._exchange(self.recv, (self.cap + 1) * RECORD, advance=True)
        self.backend.build(True)


def connect_p2p_local(flocks):
    """Ranks that live in ONE process (tests): wire their inboxes with direct device pointers.  Every flock must run on its
    own CUDA stream -- a rank's step graph waits inside the device for the flags of its peers."""
    ptrs = [f.backend.p2p_alloc(f.world, f.rank, f.cap)[1] for f in flocks]
    for f in flocks:
        f.backend.p2p_connect(direct_ptrs=ptrs)
        f.backend.p2p_seed(f.recv)


def torch_all_gather_object(world):
    """``all_gather_object`` closure on torch.distributed (default group) for ``ShardedFlock.enable_p2p``."""
    import torch.distributed as dist

    def gather(obj):
        out = [None] * world
        dist.all_gather_object(out, obj)
        return out
    return gather


def nccl_all_gather(world, cap, device):
    """all_gather closure on torch.distributed (default group); buffers are float64 CUDA tensors."""
    import torch
    import torch.distributed as dist
    recv = torch.zeros((world, cap + 1, RECORD), dtype=torch.float64, device=device)

    def gather(send):
        dist.all_gather_into_tensor(recv.view(-1), send.view(-1))
        return recv
    return gather
