"""oracle.train (numpy restatement of DAGGER.gradient_step, learner/gnn_dagger.py:76-96) against the
vectors frozen from the UNMODIFIED reference (oracle/gen_golden_train.py).  CPU only."""
import numpy as np

from conftest import rel_inf
from oracle import learner, train

TOL_LOSS = 1e-5        # relative; fp32 forward + a mean over B*N*2 squared errors
TOL_GRAD = 2e-5        # relative to the largest entry of the tensor
TOL_PARAM = 1e-6       # absolute drift allowed per Adam step, in units of ... see below


def test_first_step_loss_and_gradients(train_golden):
    g = train_golden
    layers = learner.weights_from_state_dict(g["sd0"])
    loss, grads = train.loss_and_grads(layers, g["z"][0], g["target"][0])
    assert abs(loss - g["loss"][0]) <= TOL_LOSS * abs(g["loss"][0])
    for i, (gw, gb) in enumerate(grads):
        ref_w = g["grad1"][f"conv_layers.{i}.weight"]
        ref_b = g["grad1"][f"conv_layers.{i}.bias"]
        assert rel_inf(gw.reshape(ref_w.shape), ref_w) <= TOL_GRAD
        assert rel_inf(gb, ref_b) <= TOL_GRAD


def test_adam_trajectory(train_golden):
    g = train_golden
    layers = learner.weights_from_state_dict(g["sd0"])
    opt = train.Adam(layers, lr=g["lr"])
    batches = [(g["z"][s], g["target"][s]) for s in range(g["steps"])]
    one, losses1 = train.gradient_steps(layers, opt, batches[:1])
    # a single Adam step moves every parameter by ~lr; the restatement must agree to a small fraction of that
    for i, (w, b) in enumerate(one):
        ref_w = g["sd1"][f"conv_layers.{i}.weight"]
        assert np.abs(w.reshape(ref_w.shape) - ref_w).max() <= 2e-3 * g["lr"] + 1e-7 * np.abs(ref_w).max()
        assert np.abs(b - g["sd1"][f"conv_layers.{i}.bias"]).max() <= 2e-3 * g["lr"] + 1e-7
    opt = train.Adam(layers, lr=g["lr"])
    final, losses = train.gradient_steps(layers, opt, batches)
    np.testing.assert_allclose(losses, g["loss"], rtol=5e-5)
    for i, (w, b) in enumerate(final):
        ref_w = g["sdT"][f"conv_layers.{i}.weight"]
        assert np.abs(w.reshape(ref_w.shape) - ref_w).max() <= 1e-2 * g["lr"] * g["steps"] + 1e-7 * np.abs(ref_w).max()
        assert np.abs(b - g["sdT"][f"conv_layers.{i}.bias"]).max() <= 1e-2 * g["lr"] * g["steps"] + 1e-7


def test_flatten_matches_conv_weight_order():
    rng = np.random.default_rng(0)
    z = rng.standard_normal((2, 3, 6, 5)).astype(np.float32)
    w = rng.standard_normal((4, 6, 3)).astype(np.float32)
    rows = train.flatten_inputs(z)
    direct = np.einsum('gfk,bkfn->bng', w, z).reshape(10, 4)
    np.testing.assert_allclose(rows @ w.reshape(4, 18).T, direct, rtol=1e-5, atol=1e-5)
