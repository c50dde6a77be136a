"""numpy restatement of the reference's imitation-learning update.

TEST INFRASTRUCTURE (see oracle/__init__.py).  Pinned against the reference's own
``DAGGER.gradient_step`` through tests/golden/train_*.npz (oracle/gen_golden_train.py).

Follows
* learner/gnn_dagger.py:76-96   -> ``gradient_step``: Actor.forward on the batch, F.mse_loss against the expert
                                   actions, backward, ``actor_optim.step()``
* learner/gnn_dagger.py:49      -> ``Adam(self.actor.parameters(), lr=actor_lr)`` (torch defaults: betas
                                   (0.9, 0.999), eps 1e-8, no weight decay, no amsgrad)
* learner/actor.py:73-82        -> the readout whose parameters are trained

With ind_agg = 0 (learner/gnn_dagger.py:43) the graph aggregation sits in FRONT of every trainable layer,
so no gradient flows through it: the update needs only the aggregated features z = delay_state @ delay_gso
(actor.py:70) of every sampled state, the expert label, and the MLP.
"""
import numpy as np


def flatten_inputs(z):
    """z (B,K,F,N) aggregated features -> rows (B*N, F*K) with column c = f*K + k: the order in which
    the layer-0 conv weight (H,F,K,1) is contiguous (actor.py:65,73 permutes to (B,F,K,N))."""
    B, K, F, N = z.shape
    return np.ascontiguousarray(z.transpose(0, 3, 2, 1)).reshape(B * N, F * K).astype(np.float32)


def forward(layers, rows):
    """MLP on rows (R, F*K): list of activations [a_0=rows, a_1, ..., a_L, out]."""
    acts = [rows.astype(np.float32)]
    n = len(layers)
    for i, (w, b) in enumerate(layers):
        w2 = w.reshape(w.shape[0], -1).astype(np.float32)
        x = acts[-1] @ w2.T + b.astype(np.float32)
        if i < n - 1:
            x = np.tanh(x.astype(np.float32))
        acts.append(x.astype(np.float32))
    return acts


def loss_and_grads(layers, z, target):
    """gradient_step up to ``policy_loss.backward()`` (gnn_dagger.py:83-93).

    z (B,K,F,N), target (B,1,A,N) -> (loss, [(dW like W, db), ...])."""
    B, K, F, N = z.shape
    rows = flatten_inputs(z)
    acts = forward(layers, rows)
    out = acts[-1]                                                    # (R, A)
    A = out.shape[1]
    y = np.ascontiguousarray(target.reshape(B, A, N).transpose(0, 2, 1)).reshape(B * N, A).astype(np.float32)
    diff = out - y
    loss = float(np.mean(diff.astype(np.float64) ** 2))
    d = (2.0 / diff.size) * diff                                      # dL/d out
    grads = [None] * len(layers)
    for i in range(len(layers) - 1, -1, -1):
        w, b = layers[i]
        w2 = w.reshape(w.shape[0], -1).astype(np.float32)
        grads[i] = ((d.T @ acts[i]).reshape(w.shape).astype(np.float32), d.sum(axis=0).astype(np.float32))
        if i > 0:
            d = (d @ w2) * (1.0 - acts[i] * acts[i])                  # tanh'
    return loss, grads


class Adam:
    """torch.optim.Adam, single-tensor CPU path with default hyper-parameters."""

    def __init__(self, layers, lr, beta1=0.9, beta2=0.999, eps=1e-8):
        self.lr, self.beta1, self.beta2, self.eps = lr, beta1, beta2, eps
        self.t = 0
        self.m = [(np.zeros_like(w), np.zeros_like(b)) for w, b in layers]
        self.v = [(np.zeros_like(w), np.zeros_like(b)) for w, b in layers]

    def step(self, layers, grads):
        self.t += 1
        bc1 = 1.0 - self.beta1 ** self.t
        bc2 = 1.0 - self.beta2 ** self.t
        step_size = np.float32(self.lr / bc1)
        bc2_sqrt = np.float32(np.sqrt(bc2))
        new = []
        for (w, b), (gw, gb), ms, vs in zip(layers, grads, self.m, self.v):
            outp = []
            for p, g, m, v in ((w, gw, ms[0], vs[0]), (b, gb, ms[1], vs[1])):
                m += (g - m) * np.float32(1.0 - self.beta1)
                v *= np.float32(self.beta2)
                v += np.float32(1.0 - self.beta2) * g * g
                denom = np.sqrt(v) / bc2_sqrt + np.float32(self.eps)
                outp.append((p - step_size * (m / denom)).astype(np.float32))
            new.append((outp[0], outp[1]))
        return new


def gradient_steps(layers, opt, batches):
    """Several consecutive ``gradient_step`` calls; returns (layers, [loss per step])."""
    losses = []
    for z, target in batches:
        loss, grads = loss_and_grads(layers, z, target)
        layers = opt.step(layers, grads)
        losses.append(loss)
    return layers, losses
