"""ctypes binding of libfgnn.so (include/fgnn.h) -- the host-side handle of the rollout engine.

There is NO CPU fallback: importing works anywhere (so CPU-only tests can check the ABI), but
constructing a :class:`FlockEngine` without the CUDA library or without a GPU raises.
PyTorch is used only for device memory and streams.
"""
import ctypes
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("FGNN_LIB") or os.path.join(_HERE, "libfgnn.so")      # FGNN_LIB: A/B of differently built libraries

ABI_SYMBOLS = [
    "fgnn_last_error", "fgnn_version", "fgnn_create", "fgnn_destroy", "fgnn_set_weights", "fgnn_reset",
    "fgnn_set_state", "fgnn_build_graph", "fgnn_integrate", "fgnn_env_step", "fgnn_integrate_f64", "fgnn_env_step_f64",
    "fgnn_policy", "fgnn_controller", "fgnn_controller_f64", "fgnn_step",
    "fgnn_rollout", "fgnn_actor_forward_dense", "fgnn_get_state", "fgnn_get_features", "fgnn_get_degrees",
    "fgnn_get_aggregated", "fgnn_get_action", "fgnn_export_network_dense", "fgnn_get_csr", "fgnn_get_stats",
    "fgnn_shard_configure", "fgnn_shard_local_step", "fgnn_shard_pack", "fgnn_shard_unpack", "fgnn_shard_step_begin",
    "fgnn_shard_step_end", "fgnn_shard_owned", "fgnn_profile_step", "fgnn_memcpy_sync", "fgnn_launch_count",
    "fgnn_set_agent_mask", "fgnn_set_dt", "fgnn_comm_unique_id", "fgnn_comm_init", "fgnn_shard_step",
    "fgnn_p2p_alloc", "fgnn_p2p_connect", "fgnn_p2p_seed", "fgnn_shard_step_p2p", "fgnn_shard_exchange_p2p",
    "fgnn_trainer_create", "fgnn_trainer_destroy", "fgnn_trainer_param_count", "fgnn_trainer_launch_count",
    "fgnn_trainer_step", "fgnn_actor_general_workspace", "fgnn_actor_forward_general",
]


FLAG_CSR_TAIL_ONLY = 1          # include/fgnn.h FGNN_FLAG_CSR_TAIL_ONLY


class FgnnConfig(ctypes.Structure):
    _fields_ = [
        ("n_agents", ctypes.c_int32), ("n_episodes", ctypes.c_int32), ("k", ctypes.c_int32),
        ("n_states", ctypes.c_int32), ("n_actions", ctypes.c_int32), ("hidden", ctypes.c_int32),
        ("n_layers", ctypes.c_int32), ("mean_pooling", ctypes.c_int32), ("half_accel_term", ctypes.c_int32),
        ("device", ctypes.c_int32), ("grid_dim", ctypes.c_int32), ("edge_capacity", ctypes.c_int32),
        ("readout_mode", ctypes.c_int32), ("grid_dim_y", ctypes.c_int32), ("shard_lo", ctypes.c_int32),
        ("shard_count", ctypes.c_int32), ("ghost_capacity", ctypes.c_int32), ("flags", ctypes.c_int32),
        ("comm_radius", ctypes.c_double), ("dt", ctypes.c_double), ("action_scalar", ctypes.c_double),
    ]


class FgnnStats(ctypes.Structure):
    _fields_ = [
        ("step", ctypes.c_int64), ("n_edges", ctypes.c_int64), ("overflow", ctypes.c_int32),
        ("grid_dim", ctypes.c_int32), ("n_cells", ctypes.c_int64), ("edge_capacity", ctypes.c_int64),
        ("n_ghosts", ctypes.c_int64),
    ]


class FgnnError(RuntimeError):
    pass


_lib = None


def load_library(path=None):
    """dlopen libfgnn.so and declare the prototypes of include/fgnn.h.  Raises if it is not built."""
    global _lib
    if _lib is not None and path is None:
        return _lib
    path = path or LIB_PATH
    if not os.path.exists(path):
        raise FgnnError(f"{path} not found: build it with `python -m multiagent_gnn_policies_b200.build` "
                        "(there is no CPU fallback)")
    lib = ctypes.CDLL(path)
    vp, i32, i64 = ctypes.c_void_p, ctypes.c_int32, ctypes.c_int64
    lib.fgnn_last_error.restype = ctypes.c_char_p
    lib.fgnn_last_error.argtypes = []
    lib.fgnn_version.restype = ctypes.c_int
    lib.fgnn_create.argtypes = [ctypes.POINTER(FgnnConfig), ctypes.POINTER(vp)]
    lib.fgnn_destroy.argtypes = [vp]
    lib.fgnn_set_weights.argtypes = [vp, i32, vp, vp, vp]
    lib.fgnn_reset.argtypes = [vp, vp, vp]
    lib.fgnn_set_state.argtypes = [vp, vp, vp]
    lib.fgnn_build_graph.argtypes = [vp, i32, vp]
    lib.fgnn_integrate.argtypes = [vp, vp, vp, vp]
    lib.fgnn_env_step.argtypes = [vp, vp, vp, vp]
    lib.fgnn_integrate_f64.argtypes = [vp, vp, vp, vp]
    lib.fgnn_env_step_f64.argtypes = [vp, vp, vp, vp]
    lib.fgnn_controller_f64.argtypes = [vp, i32, ctypes.c_double, vp, vp]
    lib.fgnn_policy.argtypes = [vp, vp, vp]
    lib.fgnn_controller.argtypes = [vp, i32, ctypes.c_double, vp, vp]
    lib.fgnn_step.argtypes = [vp, vp, vp, vp]
    lib.fgnn_rollout.argtypes = [vp, i32, vp, vp]
    lib.fgnn_actor_forward_dense.argtypes = [vp, i32, i32, vp, vp, vp, vp]
    lib.fgnn_actor_general_workspace.argtypes = [i32, i32, i32, i32, vp]
    lib.fgnn_actor_general_workspace.restype = i64
    lib.fgnn_actor_forward_general.argtypes = [i32, i32, i32, i32, i32, vp, i32, vp, vp, vp, vp, vp, vp, vp]
    lib.fgnn_get_state.argtypes = [vp, vp, vp]
    lib.fgnn_get_features.argtypes = [vp, i32, vp, vp]
    lib.fgnn_get_degrees.argtypes = [vp, i32, vp, vp]
    lib.fgnn_get_aggregated.argtypes = [vp, vp, vp]
    lib.fgnn_get_action.argtypes = [vp, vp, vp]
    lib.fgnn_export_network_dense.argtypes = [vp, i32, vp, vp]
    lib.fgnn_get_csr.argtypes = [vp, i32, ctypes.POINTER(vp), ctypes.POINTER(vp), ctypes.POINTER(vp), ctypes.POINTER(vp)]
    lib.fgnn_get_stats.argtypes = [vp, ctypes.POINTER(FgnnStats), vp]
    dbl = ctypes.c_double
    lib.fgnn_shard_configure.argtypes = [vp, vp, i32, i32, dbl, dbl, dbl, i32]
    lib.fgnn_shard_local_step.argtypes = [vp, vp]
    lib.fgnn_shard_pack.argtypes = [vp, vp, i64, vp, i32, i32, vp]
    lib.fgnn_shard_unpack.argtypes = [vp, vp, i32, vp]
    lib.fgnn_shard_step_begin.argtypes = [vp, vp, i64, vp, i32, vp]
    lib.fgnn_shard_step_end.argtypes = [vp, vp, i32, vp]
    lib.fgnn_shard_owned.argtypes = [vp, vp, ctypes.POINTER(i32), vp]
    lib.fgnn_profile_step.argtypes = [vp, i32, vp, vp, ctypes.POINTER(i32), vp]
    lib.fgnn_memcpy_sync.argtypes = [vp, vp, ctypes.c_uint64, vp]
    lib.fgnn_launch_count.argtypes = [vp]
    lib.fgnn_launch_count.restype = i64
    lib.fgnn_comm_unique_id.argtypes = [vp]
    lib.fgnn_comm_init.argtypes = [vp, vp, i32, i32]
    lib.fgnn_shard_step.argtypes = [vp, vp, vp, i32, vp]
    lib.fgnn_p2p_alloc.argtypes = [vp, i32, i32, i32, vp, ctypes.POINTER(vp)]
    lib.fgnn_p2p_connect.argtypes = [vp, vp, vp]
    lib.fgnn_p2p_seed.argtypes = [vp, vp, vp]
    lib.fgnn_shard_step_p2p.argtypes = [vp, vp]
    lib.fgnn_shard_exchange_p2p.argtypes = [vp, i32, vp]
    lib.fgnn_set_agent_mask.argtypes = [vp, vp, vp]
    lib.fgnn_set_dt.argtypes = [vp, dbl]
    lib.fgnn_trainer_create.argtypes = [i32, i32, i32, i32, ctypes.POINTER(vp)]
    lib.fgnn_trainer_destroy.argtypes = [vp]
    lib.fgnn_trainer_param_count.argtypes = [vp]
    lib.fgnn_trainer_param_count.restype = i32
    lib.fgnn_trainer_launch_count.argtypes = [vp]
    lib.fgnn_trainer_launch_count.restype = i64
    lib.fgnn_trainer_step.argtypes = [vp, i32, i32, vp, vp, vp, vp, vp, i64, dbl, dbl, dbl, dbl, i32, vp, vp, vp]
    for name in ABI_SYMBOLS:
        if name not in ("fgnn_last_error", "fgnn_launch_count", "fgnn_version", "fgnn_trainer_param_count",
                        "fgnn_trainer_launch_count"):
            getattr(lib, name).restype = ctypes.c_int
    if path == LIB_PATH:
        _lib = lib
    return lib


def _ptr(obj):
    """Raw address of a numpy array, a torch tensor, an int, or None."""
    if obj is None:
        return None
    if isinstance(obj, int):
        return obj
    if isinstance(obj, np.ndarray):
        assert obj.flags["C_CONTIGUOUS"], "array must be C-contiguous"
        return obj.ctypes.data
    if hasattr(obj, "data_ptr"):
        assert obj.is_contiguous(), "tensor must be contiguous"
        return obj.data_ptr()
    raise TypeError(f"cannot take the address of {type(obj)}")


class FlockEngine:
    """One engine handle = B episodes x N agents of FlockingRelative-v0 + a K-tap aggregation-GNN actor.

    Mirrors, on device, what the reference spreads over env.step (gym_flock),
    MultiAgentStateWithDelay (learner/state_with_delay.py:6-53) and DAGGER.select_action
    (learner/gnn_dagger.py:55-72)."""

    def __init__(self, n_agents, k=3, hidden=32, n_layers=2, comm_radius=1.0, dt=0.01, n_episodes=1,
                 device=0, action_scalar=10.0, mean_pooling=True, half_accel_term=True, grid_dim=0,
                 edge_capacity=0, readout_mode=0, n_states=6, n_actions=2, stream=None, grid_dim_y=0,
                 shard_lo=0, shard_count=0, ghost_capacity=0, csr_tail_only=False):
        import torch  # device memory + streams only
        if not torch.cuda.is_available():
            raise FgnnError("no CUDA device: the rollout engine has no CPU fallback")
        self._torch = torch
        self.lib = load_library()
        self.n_agents, self.n_episodes, self.k = int(n_agents), int(n_episodes), int(k)
        self.hidden, self.n_layers = int(hidden), int(n_layers)
        self.n_states, self.n_actions = int(n_states), int(n_actions)
        self.comm_radius, self.dt, self.action_scalar = float(comm_radius), float(dt), float(action_scalar)
        self.device_index = int(device)
        self.device = torch.device("cuda", self.device_index)
        self.M = self.n_agents * self.n_episodes
        self._stream = stream
        cfg = FgnnConfig(self.n_agents, self.n_episodes, self.k, self.n_states, self.n_actions, self.hidden,
                         self.n_layers, int(bool(mean_pooling)), int(bool(half_accel_term)), self.device_index,
                         int(grid_dim), int(edge_capacity), int(readout_mode), int(grid_dim_y), int(shard_lo),
                         int(shard_count), int(ghost_capacity), FLAG_CSR_TAIL_ONLY if csr_tail_only else 0,
                         self.comm_radius, self.dt, self.action_scalar)
        self.shard_lo, self.shard_count, self.ghost_capacity = int(shard_lo), int(shard_count), int(ghost_capacity)
        # rows of the action arrays policy()/integrate() exchange: all agents, or (sharded) the list capacity
        self.rows_io = (self.shard_count + self.ghost_capacity) if self.shard_count else self.M
        self._h = ctypes.c_void_p()
        self._check(self.lib.fgnn_create(ctypes.byref(cfg), ctypes.byref(self._h)))
        self.step_index = -1          # host mirror of the engine's step counter t (-1: never reset)

    # -- plumbing ---------------------------------------------------------------------------
    def _check(self, rc):
        if rc != 0:
            raise FgnnError(self.lib.fgnn_last_error().decode())

    @property
    def stream(self):
        if self._stream is not None:
            return self._stream
        return self._torch.cuda.current_stream(self.device).cuda_stream

    def sync(self):
        """Wait for the stream the engine enqueues on (an explicit ``stream=`` handle, else torch's current stream)."""
        if self._stream is not None:
            self._torch.cuda.ExternalStream(self._stream, device=self.device).synchronize()
        else:
            self._torch.cuda.current_stream(self.device).synchronize()

    def close(self):
        if getattr(self, "_h", None) is not None and self._h:
            self.lib.fgnn_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    # -- weights ----------------------------------------------------------------------------
    def load_state_dict(self, sd):
        """``conv_layers.{i}.weight`` (out,in,step,1) / ``.bias`` as in the reference checkpoint."""
        for i in range(self.n_layers + 1):
            w = sd[f"conv_layers.{i}.weight"]
            b = sd[f"conv_layers.{i}.bias"]
            w = np.ascontiguousarray(w.detach().cpu().numpy() if hasattr(w, "detach") else w, dtype=np.float32)
            b = np.ascontiguousarray(b.detach().cpu().numpy() if hasattr(b, "detach") else b, dtype=np.float32)
            out_dim = self.n_actions if i == self.n_layers else self.hidden
            in_dim = self.n_states if i == 0 else self.hidden
            step = self.k if i == 0 else 1
            if w.reshape(w.shape[0], w.shape[1], -1).shape != (out_dim, in_dim, step):
                raise FgnnError(f"layer {i}: weight shape {w.shape} does not match ({out_dim},{in_dim},{step},1)")
            self._check(self.lib.fgnn_set_weights(self._h, i, _ptr(w), _ptr(b), self.stream))

    # -- env side ---------------------------------------------------------------------------
    def _as_state(self, x):
        if hasattr(x, "data_ptr"):
            assert x.dtype == self._torch.float64 and x.numel() == self.M * 4
            return x.contiguous()
        x = np.ascontiguousarray(x, dtype=np.float64)
        assert x.size == self.M * 4, f"state must have {self.M}x4 entries"
        return x

    def reset(self, x):
        x = self._as_state(x)
        self._check(self.lib.fgnn_reset(self._h, _ptr(x), self.stream))
        self.step_index = 0
        if isinstance(x, np.ndarray):
            self.sync()

    def set_state(self, x):
        x = self._as_state(x)
        self._check(self.lib.fgnn_set_state(self._h, _ptr(x), self.stream))
        if isinstance(x, np.ndarray):
            self.sync()

    def build_graph(self, advance=True):
        self._check(self.lib.fgnn_build_graph(self._h, int(bool(advance)), self.stream))
        self.step_index += int(bool(advance))

    def _as_action(self, u):
        """(array, is_f64): float64 actions (the controller's, learner/gnn_dagger.py:156-163) stay float64; everything
        else is taken as fp32 (select_action's ``action.cpu().numpy()``, learner/gnn_dagger.py:161)."""
        if hasattr(u, "data_ptr"):
            assert u.dtype in (self._torch.float32, self._torch.float64) and u.numel() == self.rows_io * 2
            return u.contiguous(), u.dtype == self._torch.float64
        u = np.asarray(u)
        f64 = u.dtype == np.float64
        u = np.ascontiguousarray(u, dtype=np.float64 if f64 else np.float32)
        assert u.size == self.rows_io * 2
        return u, f64

    def integrate(self, u, want_reward=False):
        u, f64 = self._as_action(u)
        r = np.empty(self.n_episodes, dtype=np.float64) if want_reward else None
        fn = self.lib.fgnn_integrate_f64 if f64 else self.lib.fgnn_integrate
        self._check(fn(self._h, _ptr(u), _ptr(r), self.stream))
        if isinstance(u, np.ndarray) or want_reward:
            self.sync()
        return r

    def env_step(self, u):
        """env.step(u): integrate + rebuild graph/features; returns the per-episode reward (B,) f64."""
        u, f64 = self._as_action(u)
        r = np.empty(self.n_episodes, dtype=np.float64)
        fn = self.lib.fgnn_env_step_f64 if f64 else self.lib.fgnn_env_step
        self._check(fn(self._h, _ptr(u), _ptr(r), self.stream))
        self.step_index += 1
        self.sync()
        return r

    # -- learner side -----------------------------------------------------------------------
    def policy(self, out=None):
        """select_action: (B*N, 2) fp32.  ``out`` may be a torch CUDA tensor or a (pinned) numpy array;
        default is a fresh CUDA tensor."""
        if out is None:
            out = self._torch.empty((self.rows_io, 2), dtype=self._torch.float32, device=self.device)
        self._check(self.lib.fgnn_policy(self._h, _ptr(out), self.stream))
        if isinstance(out, np.ndarray):
            self.sync()
        return out

    def controller(self, centralized=True, max_accel=1.0, out=None, dtype=np.float32):
        """Expert controller action (B*N,2) for the current state (env.env.controller): fp32, or float64 as gym_flock
        returns it (``dtype=np.float64`` or a float64 ``out``)."""
        if out is None:
            out = np.empty((self.M, 2), dtype=dtype)
        f64 = (out.dtype == np.float64) if isinstance(out, np.ndarray) else (out.dtype == self._torch.float64)
        fn = self.lib.fgnn_controller_f64 if f64 else self.lib.fgnn_controller
        self._check(fn(self._h, int(bool(centralized)), float(max_accel), _ptr(out), self.stream))
        if isinstance(out, np.ndarray):
            self.sync()
        return out

    def step(self, action_out=None, reward_out=None):
        """One closed-loop step (select_action -> env.step).  Outputs optional, see ``policy``."""
        self._check(self.lib.fgnn_step(self._h, _ptr(action_out), _ptr(reward_out), self.stream))
        self.step_index += 1
        if isinstance(action_out, np.ndarray) or isinstance(reward_out, np.ndarray):
            self.sync()

    def rollout(self, steps, want_reward=False):
        r = np.empty((steps, self.n_episodes), dtype=np.float64) if want_reward else None
        self._check(self.lib.fgnn_rollout(self._h, int(steps), _ptr(r), self.stream))
        self.step_index += int(steps)
        if want_reward:
            self.sync()
        return r

    def actor_forward_dense(self, delay_state, delay_gso):
        """Actor.forward on dense CUDA tensors: (B,K,F,N), (B,K,N,N) -> (B,1,A,N)."""
        torch = self._torch
        ds = delay_state.to(self.device, torch.float32).contiguous()
        gso = delay_gso.to(self.device, torch.float32).contiguous()
        B, K, Fdim, N = ds.shape
        assert K == self.k and Fdim == self.n_states and tuple(gso.shape) == (B, K, N, N)
        out = torch.empty((B, 1, self.n_actions, N), dtype=torch.float32, device=self.device)
        self._check(self.lib.fgnn_actor_forward_dense(self._h, B, N, _ptr(ds), _ptr(gso), _ptr(out), self.stream))
        return out

    # -- read-back --------------------------------------------------------------------------
    def get_state(self):
        x = np.empty((self.M, 4), dtype=np.float64)
        self._check(self.lib.fgnn_get_state(self._h, _ptr(x), self.stream))
        self.sync()
        return x

    def get_features(self, age=0):
        v = np.empty((self.M, 6), dtype=np.float32)
        self._check(self.lib.fgnn_get_features(self._h, age, _ptr(v), self.stream))
        self.sync()
        return v

    def get_degrees(self, age=0):
        d = np.empty(self.M, dtype=np.int32)
        self._check(self.lib.fgnn_get_degrees(self._h, age, _ptr(d), self.stream))
        self.sync()
        return d

    def get_aggregated(self, device=False):
        """(K, B*N, 6) aggregated features z_k of the last policy call (numpy, or a CUDA tensor)."""
        if device:
            z = self._torch.empty((self.k, self.M, 6), dtype=self._torch.float32, device=self.device)
            self._check(self.lib.fgnn_get_aggregated(self._h, _ptr(z), self.stream))
            return z
        z = np.empty((self.k, self.M, 6), dtype=np.float32)
        self._check(self.lib.fgnn_get_aggregated(self._h, _ptr(z), self.stream))
        self.sync()
        return z

    def aggregate(self):
        """K-hop aggregation of the current history (actor.py:68-71 on the sparse graphs): (K, B*N, 6) CUDA
        tensor.  This is everything ``gradient_step`` needs from a state (ind_agg = 0)."""
        self._check(self.lib.fgnn_policy(self._h, None, self.stream))
        return self.get_aggregated(device=True)

    # -- env variants ---------------------------------------------------
    def set_agent_mask(self, mask):
        """FlockingLeader: ``mask`` (B*N,) with 0 for leaders (their action is ignored); None removes it."""
        if mask is not None:
            mask = np.ascontiguousarray(np.asarray(mask) != 0, dtype=np.uint8)
            assert mask.size == self.M
        self._check(self.lib.fgnn_set_agent_mask(self._h, _ptr(mask), self.stream))
        self.sync()

    def set_dt(self, dt):
        """FlockingStochastic: the time step of the following integrations."""
        self._check(self.lib.fgnn_set_dt(self._h, float(dt)))
        self.dt = float(dt)

    def get_action(self):
        a = np.empty((self.M, 2), dtype=np.float32)
        self._check(self.lib.fgnn_get_action(self._h, _ptr(a), self.stream))
        self.sync()
        return a

    def network_dense(self, age=0, device=False):
        """Row-normalised state_network of graph t-age as (B,N,N) fp32 (numpy, or CUDA tensor)."""
        shape = (self.n_episodes, self.n_agents, self.n_agents)
        if device:
            out = self._torch.empty(shape, dtype=self._torch.float32, device=self.device)
            self._check(self.lib.fgnn_export_network_dense(self._h, age, _ptr(out), self.stream))
            return out
        out = np.empty(shape, dtype=np.float32)
        self._check(self.lib.fgnn_export_network_dense(self._h, age, _ptr(out), self.stream))
        self.sync()
        return out

    def csr(self, age=0):
        """(row_start uint32 (M,), deg int32 (M,), cols int32 (nnz_cap,), scale fp32 (M,)) as numpy copies."""
        rs, dg, cl, sc = (ctypes.c_void_p() for _ in range(4))
        self._check(self.lib.fgnn_get_csr(self._h, age, ctypes.byref(rs), ctypes.byref(dg), ctypes.byref(cl),
                                          ctypes.byref(sc)))

        def fetch(ptr, n, dtype):
            host = np.empty(n, dtype=dtype)
            self._check(self.lib.fgnn_memcpy_sync(host.ctypes.data, ptr, host.nbytes, self.stream))
            return host
        deg = fetch(dg, self.M, np.int32)
        row_start = fetch(rs, self.M, np.uint32)
        n_used = int((row_start.astype(np.int64) + deg).max()) if self.M else 0
        cols = fetch(cl, max(n_used, 1), np.int32)[:n_used]
        scale = fetch(sc, self.M, np.float32)
        return row_start, deg, cols, scale

    def stats(self):
        s = FgnnStats()
        self._check(self.lib.fgnn_get_stats(self._h, ctypes.byref(s), self.stream))
        return {"step": s.step, "n_edges": s.n_edges, "overflow": int(s.overflow), "grid_dim": s.grid_dim,
                "n_cells": s.n_cells, "edge_capacity": s.edge_capacity, "n_ghosts": s.n_ghosts}

    # -- multi-GPU pieces (orchestrated by parallel.ShardedFlock) ---------------------------
    def shard_configure(self, bounds, world, rank, depth, margin, dshift, handover_after):
        bounds = np.ascontiguousarray(bounds, dtype=np.float64)
        assert bounds.size == world + 1
        self._check(self.lib.fgnn_shard_configure(self._h, _ptr(bounds), world, rank, float(depth), float(margin),
                                                  float(dshift), int(handover_after)))

    def shard_local_step(self):
        self._check(self.lib.fgnn_shard_local_step(self._h, self.stream))

    def shard_pack(self, windows, window_stride, send_buf, cap, advance):
        self._check(self.lib.fgnn_shard_pack(self._h, _ptr(windows), int(window_stride), _ptr(send_buf), cap,
                                             int(bool(advance)), self.stream))

    def shard_unpack(self, recv_buf, cap):
        self._check(self.lib.fgnn_shard_unpack(self._h, _ptr(recv_buf), cap, self.stream))

    def shard_step_begin(self, windows, window_stride, send_buf, cap):
        self._check(self.lib.fgnn_shard_step_begin(self._h, _ptr(windows), int(window_stride), _ptr(send_buf), cap,
                                                   self.stream))

    def shard_step_end(self, recv_buf, cap):
        self._check(self.lib.fgnn_shard_step_end(self._h, _ptr(recv_buf), cap, self.stream))
        self.step_index += 1

    def comm_init(self, unique_id, rank, world):
        """Collective: create this rank's NCCL communicator from the 128-byte id rank 0 made (``comm_unique_id``)."""
        buf = ctypes.create_string_buffer(bytes(unique_id), 128)
        self._check(self.lib.fgnn_comm_init(self._h, ctypes.addressof(buf), int(rank), int(world)))

    def shard_step(self, send_buf, recv_buf, cap):
        """The whole sharded step as one CUDA graph with the NCCL all-gather inside."""
        self._check(self.lib.fgnn_shard_step(self._h, _ptr(send_buf), _ptr(recv_buf), cap, self.stream))
        self.step_index += 1

    # p2p halo transport: records stored straight into the peers' inboxes (CUDA-IPC over NVLink), one graph per step
    def p2p_alloc(self, world, rank, cap):
        """Allocate this rank's inbox; returns (64-byte cudaIpcMemHandle_t as bytes, local device pointer)."""
        handle = ctypes.create_string_buffer(64)
        ptr = ctypes.c_void_p()
        self._check(self.lib.fgnn_p2p_alloc(self._h, int(world), int(rank), int(cap), ctypes.addressof(handle), ctypes.byref(ptr)))
        return handle.raw, ptr.value

    def p2p_connect(self, handles=None, direct_ptrs=None):
        """``handles``: the ranks' IPC handles concatenated in rank order (separate processes); or ``direct_ptrs``: the
        inbox device pointers of ranks living in this process."""
        if direct_ptrs is not None:
            arr = (ctypes.c_void_p * len(direct_ptrs))(*direct_ptrs)
            self._check(self.lib.fgnn_p2p_connect(self._h, None, ctypes.addressof(arr)))
        else:
            buf = ctypes.create_string_buffer(bytes(handles), len(handles))
            self._check(self.lib.fgnn_p2p_connect(self._h, ctypes.addressof(buf), None))

    def p2p_seed(self, gathered):
        self._check(self.lib.fgnn_p2p_seed(self._h, _ptr(gathered), self.stream))

    def shard_step_p2p(self):
        self._check(self.lib.fgnn_shard_step_p2p(self._h, self.stream))
        self.step_index += 1

    def shard_exchange_p2p(self, advance=True):
        self._check(self.lib.fgnn_shard_exchange_p2p(self._h, int(bool(advance)), self.stream))

    def shard_owned(self):
        """Currently owned agents (global ids, int32, list order) -- synchronises."""
        ids = np.empty(self.shard_count + self.ghost_capacity, dtype=np.int32)
        n = ctypes.c_int32(0)
        self._check(self.lib.fgnn_shard_owned(self._h, _ptr(ids), ctypes.byref(n), self.stream))
        ids = ids[:n.value]
        return ids[ids >= 0].copy()          # -1 = slot of an agent that was handed over

    def profile_step(self):
        """One closed-loop step with per-kernel CUDA-event timing: [(kernel name, ms), ...]."""
        ms = (ctypes.c_float * 16)()
        names = ctypes.create_string_buffer(16 * 16)
        n = ctypes.c_int32(0)
        self._check(self.lib.fgnn_profile_step(self._h, 16, ctypes.addressof(ms), ctypes.addressof(names),
                                               ctypes.byref(n), self.stream))
        self.step_index += 1
        return [(names.raw[16 * i:16 * i + 16].split(b"\0")[0].decode(), float(ms[i])) for i in range(n.value)]

    def launch_count(self):
        return int(self.lib.fgnn_launch_count(self._h))


def comm_unique_id():
    """128-byte NCCL id (call on rank 0, broadcast to the other ranks, pass to ``FlockEngine.comm_init``)."""
    lib = load_library()
    buf = ctypes.create_string_buffer(128)
    if lib.fgnn_comm_unique_id(ctypes.addressof(buf)) != 0:
        raise FgnnError(lib.fgnn_last_error().decode())
    return buf.raw


class ActorTrainer:
    """Native ``gradient_step`` (learner/gnn_dagger.py:76-96): MLP forward, MSE loss, backward and Adam on the
    device, updating the torch parameters and the torch.optim.Adam state tensors IN PLACE.

    ``params`` / ``exp_avg`` / ``exp_avg_sq`` are lists of CUDA tensors in the order W_0, b_0, ..., W_L, b_L
    (conv layout of learner/actor.py:30-40)."""

    def __init__(self, k, hidden, n_layers, device=0):
        import torch
        if not torch.cuda.is_available():
            raise FgnnError("no CUDA device: the trainer has no CPU fallback")
        self._torch = torch
        self.lib = load_library()
        self.k, self.hidden, self.n_layers = int(k), int(hidden), int(n_layers)
        self.device = torch.device("cuda", int(device))
        self._h = ctypes.c_void_p()
        self._check(self.lib.fgnn_trainer_create(self.k, self.hidden, self.n_layers, int(device), ctypes.byref(self._h)))
        self.n_params = int(self.lib.fgnn_trainer_param_count(self._h))

    def _check(self, rc):
        if rc != 0:
            raise FgnnError(self.lib.fgnn_last_error().decode())

    def close(self):
        if getattr(self, "_h", None) is not None and self._h:
            self.lib.fgnn_trainer_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def launch_count(self):
        return int(self.lib.fgnn_trainer_launch_count(self._h))

    def _ptr_array(self, tensors):
        n = 2 * (self.n_layers + 1)
        assert len(tensors) == n, f"expected {n} tensors (W_0, b_0, ..., W_L, b_L)"
        for t in tensors:
            assert t.is_cuda and t.dtype == self._torch.float32 and t.is_contiguous()
        return (ctypes.c_void_p * n)(*[t.data_ptr() for t in tensors])

    def step(self, z, target, params, exp_avg=None, exp_avg_sq=None, step=1, lr=1e-3, betas=(0.9, 0.999), eps=1e-8,
             apply=True, want_grads=False):
        """z (B,K,N,6), target (B,1,2,N) or (B,2,N): CUDA fp32.  Returns (loss tensor (1,), grads tensor or None)."""
        torch = self._torch
        z = z.to(self.device, torch.float32).contiguous()
        target = target.to(self.device, torch.float32).contiguous()
        B, K, N, Fd = z.shape
        assert K == self.k and Fd == 6 and target.numel() == B * 2 * N
        loss = torch.empty(1, dtype=torch.float32, device=self.device)
        grads = torch.empty(self.n_params, dtype=torch.float32, device=self.device) if want_grads else None
        pp = self._ptr_array(params)
        mm = self._ptr_array(exp_avg) if apply else None
        vv = self._ptr_array(exp_avg_sq) if apply else None
        stream = torch.cuda.current_stream(self.device).cuda_stream
        self._check(self.lib.fgnn_trainer_step(self._h, B, N, _ptr(z), _ptr(target), pp, mm, vv, int(step), float(lr),
                                               float(betas[0]), float(betas[1]), float(eps), int(bool(apply)),
                                               _ptr(loss), _ptr(grads), stream))
        return loss, grads
