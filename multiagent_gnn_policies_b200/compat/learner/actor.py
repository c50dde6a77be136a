"""Aggregation-GNN policy with the reference's constructor, attributes, parameter names and
``forward(delay_state, delay_gso)`` contract (learner/actor.py:7-86).

Inference (no autograd) runs in libfgnn.so: the dense-tensor kernel ``fgnn_actor_forward_dense``
(per-agent K-tap aggregation + readout on CUDA cores / tensor cores) for DAGGER's shape (ind_agg = 0, equal hidden
widths), ``fgnn_actor_forward_general`` for everything else the constructor accepts (ind_agg > 0 as in
learner/gnn_ddpg.py:126, unequal widths, n_s != 6).  When autograd is recording
(DAGGER.gradient_step, learner/gnn_dagger.py:76-96) the same arithmetic is expressed with torch ops on
the GPU so gradients flow -- training is a "next" row of SURVEY.md section 8(f), not the hot path.
There is no CPU path: tensors must live on a CUDA device.
"""
import ctypes

import torch
import torch.nn as nn

from multiagent_gnn_policies_b200.engine import FlockEngine, FgnnError, load_library


class Actor(nn.Module):

    def __init__(self, n_s, n_a, hidden_layers, k, ind_agg):
        super(Actor, self).__init__()
        self.k = k
        self.n_s = n_s
        self.n_a = n_a
        self.layers = [n_s] + list(hidden_layers) + [n_a]
        self.n_layers = len(self.layers) - 1
        self.ind_agg = ind_agg
        convs = []
        for i in range(self.n_layers):
            step = k if i == ind_agg else 1
            convs.append(nn.Conv2d(in_channels=self.layers[i], out_channels=self.layers[i + 1],
                                   kernel_size=(step, 1), stride=(step, 1)))
        self.conv_layers = nn.ModuleList(convs)
        self._engine = None
        self._engine_versions = None
        self.native_updates = 0        # bumped by the native gradient step (it bypasses torch's version counters)

    # -- engine plumbing --------------------------------------------------------------------
    def _engine_supported(self):
        hidden = self.layers[1:-1]
        return (self.ind_agg == 0 and self.n_s == 6 and self.n_a == 2 and 1 <= len(hidden) <= 4
                and len(set(hidden)) == 1 and hidden[0] <= 128 and 1 <= self.k <= 4)

    def weight_versions(self):
        return (self.native_updates,) + tuple(p._version for p in self.parameters())

    def sync_engine(self, engine):
        """Push the current parameters into ``engine`` if they changed since the last push."""
        versions = (id(engine),) + self.weight_versions()
        if getattr(engine, "_actor_versions", None) != versions:
            engine.load_state_dict({k: v.detach() for k, v in self.state_dict().items()})
            engine._actor_versions = versions

    def _dense_engine(self, device):
        if self._engine is None or self._engine.device != device:
            self._engine = FlockEngine(n_agents=1, k=self.k, hidden=self.layers[1], n_layers=self.n_layers - 1,
                                       device=device.index if device.index is not None else torch.cuda.current_device())
        self.sync_engine(self._engine)
        return self._engine

    def _forward_general(self, delay_state, delay_gso):
        """learner/actor.py:45-86 for any ``ind_agg`` / layer widths: three plain CUDA kernels behind
        ``fgnn_actor_forward_general`` (csrc/fgnn_dense.cu), parameters read in place from the conv tensors."""
        lib = load_library()
        dev = delay_state.device
        B, K, _, N = delay_state.shape
        L = self.n_layers
        ds = delay_state.detach().to(torch.float32).contiguous()
        gso = delay_gso.detach().to(device=dev, dtype=torch.float32).contiguous()
        ws = [c.weight.detach().to(torch.float32).contiguous() for c in self.conv_layers]
        bs = [c.bias.detach().to(torch.float32).contiguous() for c in self.conv_layers]
        if any(t.device != dev for t in ws + bs):
            raise FgnnError(f"Actor.forward: parameters on {ws[0].device}, inputs on {dev}")
        widths = (ctypes.c_int32 * (L + 1))(*self.layers)
        need = lib.fgnn_actor_general_workspace(B, N, K, L, widths)
        if need < 0:
            raise FgnnError(lib.fgnn_last_error().decode())
        work = torch.empty((need // 4,), dtype=torch.float32, device=dev)
        out = torch.empty((B, 1, self.n_a, N), dtype=torch.float32, device=dev)
        wp = (ctypes.c_void_p * L)(*[t.data_ptr() for t in ws])
        bp = (ctypes.c_void_p * L)(*[t.data_ptr() for t in bs])
        rc = lib.fgnn_actor_forward_general(dev.index, B, N, K, L, widths, self.ind_agg, wp, bp, ds.data_ptr(), gso.data_ptr(),
                                            out.data_ptr(), work.data_ptr(), torch.cuda.current_stream(dev).cuda_stream)
        if rc != 0:
            raise FgnnError(lib.fgnn_last_error().decode())
        return out

    # -- forward ----------------------------------------------------------------------------
    def forward(self, delay_state, delay_gso):
        batch_size = delay_state.shape[0]
        n_agents = delay_state.shape[3]
        assert delay_gso.shape[0] == batch_size
        assert delay_gso.shape[2] == n_agents
        assert delay_gso.shape[3] == n_agents
        assert delay_state.shape[1] == self.k
        assert delay_state.shape[2] == self.n_s
        assert delay_gso.shape[1] == self.k
        if not delay_state.is_cuda:
            raise FgnnError("Actor.forward needs CUDA tensors: this build has no CPU fallback")
        needs_grad = torch.is_grad_enabled() and (any(p.requires_grad for p in self.parameters())
                                                  or delay_state.requires_grad)
        if not needs_grad and self._engine_supported():
            eng = self._dense_engine(delay_state.device)
            return eng.actor_forward_dense(delay_state, delay_gso)
        if not needs_grad and self.n_layers <= 16 and (0 <= self.ind_agg < self.n_layers or self.k == 1):
            return self._forward_general(delay_state, delay_gso)
        # autograd path (training on dense tensors): same arithmetic, torch ops on the GPU so that gradients flow.  TF32 is
        # switched off for it (cuDNN convolutions default to TF32): the reference computes in fp32 on the CPU.
        matmul_tf32 = torch.backends.cuda.matmul.allow_tf32
        torch.backends.cuda.matmul.allow_tf32 = False
        try:
            with torch.backends.cudnn.flags(enabled=True, allow_tf32=False):
                x = delay_state.permute(0, 2, 1, 3)                        # (B,F,K,N)
                for i in range(self.n_layers):
                    if i == self.ind_agg:
                        x = torch.matmul(x.permute(0, 2, 1, 3), delay_gso).permute(0, 2, 1, 3)
                    x = self.conv_layers[i](x)
                    if i < self.n_layers - 1:
                        x = torch.tanh(x)
                return x.view((batch_size, 1, self.n_a, n_agents))
        finally:
            torch.backends.cuda.matmul.allow_tf32 = matmul_tf32
