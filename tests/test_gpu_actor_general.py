"""Actor.forward for aggregation indices other than DAGGER's 0 and for unequal layer widths (learner/actor.py:45-86,
what learner/gnn_ddpg.py:126 builds) through the compat ``learner.actor.Actor`` -> ``fgnn_actor_forward_general``
(csrc/fgnn_dense.cu), against fixtures made by the UNMODIFIED reference Actor (oracle/gen_golden_agg.py)."""
import numpy as np
import pytest

from conftest import rel_inf, agg_golden_names, load_agg_golden
from oracle import learner as olearner

pytestmark = pytest.mark.gpu

TOL_ACTION = 1e-5       # north_star: actions within 1e-5 relative (inf-norm) fp32


@pytest.fixture()
def compat():
    from multiagent_gnn_policies_b200 import compat as c
    c.install()
    return c


@pytest.mark.parametrize("name", agg_golden_names())
def test_general_forward_matches_reference(name, compat, monkeypatch):
    import torch
    from learner.actor import Actor
    from multiagent_gnn_policies_b200.engine import load_library
    g = load_agg_golden(name)
    layers = g["layers"]
    actor = Actor(layers[0], layers[-1], layers[1:-1], g["k"], g["ind_agg"]).to("cuda")
    actor.load_state_dict({k: torch.from_numpy(v) for k, v in g["state_dict"].items()})
    ds = torch.from_numpy(g["delay_state"]).cuda()
    gso = torch.from_numpy(g["delay_gso"]).cuda()
    # the native path must be the one that runs: count calls into the C ABI
    lib = load_library()
    calls = []
    real = lib.fgnn_actor_forward_general

    def counted(*a):
        calls.append(1)
        return real(*a)
    monkeypatch.setattr(lib, "fgnn_actor_forward_general", counted)
    assert not actor._engine_supported()
    with torch.no_grad():
        out = actor(ds, gso)
    torch.cuda.synchronize()
    assert calls, "Actor.forward did not take fgnn_actor_forward_general"
    assert out.shape == g["out"].shape and out.dtype == torch.float32 and out.is_cuda
    out = out.cpu().numpy()
    assert rel_inf(out, g["out"]) <= TOL_ACTION, rel_inf(out, g["out"])
    # ... and the oracle restatement of the same formula
    ref = olearner.actor_forward(olearner.weights_from_state_dict(g["state_dict"]), g["delay_state"], g["delay_gso"],
                                 ind_agg=g["ind_agg"])
    assert rel_inf(out, ref) <= TOL_ACTION
    # the autograd path (torch ops on the GPU, used when gradients are recorded) computes the same function
    with torch.enable_grad():
        out_grad = actor(ds, gso)
    assert out_grad.requires_grad
    assert rel_inf(out_grad.detach().cpu().numpy(), g["out"]) <= TOL_ACTION
    # a second call (fresh workspace, same stream) and a non-contiguous input view give the same bits
    with torch.no_grad():
        again = actor(ds, gso).cpu().numpy()
        ds_nc = ds.permute(0, 2, 1, 3).contiguous().permute(0, 2, 1, 3)
        nc = actor(ds_nc, gso).cpu().numpy()
    np.testing.assert_array_equal(again, out)
    np.testing.assert_array_equal(nc, out)


def test_general_forward_k1_without_aggregation(compat):
    """K = 1 with ind_agg beyond the layers: the reference never aggregates and its final view still works (one row)."""
    import torch
    from learner.actor import Actor
    torch.manual_seed(3)
    actor = Actor(6, 2, [10, 14], 1, 7).to("cuda")
    ds = torch.randn(2, 1, 6, 17, device="cuda")
    gso = torch.eye(17, device="cuda").reshape(1, 1, 17, 17).repeat(2, 1, 1, 1)
    with torch.no_grad():
        out = actor(ds, gso)
    with torch.enable_grad():
        ref = actor(ds, gso).detach()
    assert rel_inf(out.cpu().numpy(), ref.cpu().numpy()) <= TOL_ACTION
