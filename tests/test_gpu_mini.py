"""Single-CTA path for small flocks (B*N <= 128 agents, fgnn_mini.cu): the same device code as the general step behind one
launch, so fgnn_step / fgnn_rollout must leave the bits of the general kernels (FGNN_MINI=0), mixed freely with the API-split
calls (policy / env_step), for every K, several hidden widths, batched episodes and the leader mask."""
import numpy as np
import pytest

from conftest import load_golden
from oracle import flock_env

pytestmark = pytest.mark.gpu


def _random_sd(rng, k, hidden, n_layers=2):
    sd = {"conv_layers.0.weight": rng.normal(0, 0.3, (hidden, 6, k, 1)).astype(np.float32),
          "conv_layers.0.bias": rng.normal(0, 0.1, hidden).astype(np.float32)}
    for l in range(1, n_layers):
        sd[f"conv_layers.{l}.weight"] = rng.normal(0, 0.3, (hidden, hidden, 1, 1)).astype(np.float32)
        sd[f"conv_layers.{l}.bias"] = rng.normal(0, 0.1, hidden).astype(np.float32)
    sd[f"conv_layers.{n_layers}.weight"] = rng.normal(0, 0.3, (2, hidden, 1, 1)).astype(np.float32)
    sd[f"conv_layers.{n_layers}.bias"] = rng.normal(0, 0.1, 2).astype(np.float32)
    return sd


def _run(monkeypatch, mini, n, episodes, k, hidden, sd, x0, mask=None):
    from multiagent_gnn_policies_b200.engine import FlockEngine
    monkeypatch.setenv("FGNN_MINI", "1" if mini else "0")
    eng = FlockEngine(n_agents=n, n_episodes=episodes, k=k, hidden=hidden, n_layers=2, comm_radius=1.0, dt=0.01)
    eng.load_state_dict(sd)
    if mask is not None:
        eng.set_agent_mask(mask)
    eng.reset(x0)
    out = []
    m = n * episodes
    for t in range(5):                                   # fused steps
        a, r = np.empty((m, 2), np.float32), np.empty(episodes)
        eng.step(a, r)
        out.append((a.copy(), r.copy(), eng.get_state(), eng.get_degrees(), eng.get_features()))
    rew = eng.rollout(7, want_reward=True)
    out.append((eng.get_state(), eng.get_degrees(), eng.get_features(), rew))
    a = np.empty((m, 2), np.float32)                     # API-split calls after the fused ones ...
    eng.policy(out=a)
    out.append((a.copy(), eng.get_aggregated()))
    eng.env_step(a)
    eng.policy(out=a)
    out.append((a.copy(),))
    eng.env_step(a.astype(np.float64) * 0.5)             # float64 action (the controller's dtype)
    out.append((eng.get_state(), eng.get_degrees(), eng.get_features()))
    eng.rollout(3)                                       # ... and fused ones after those
    out.append((eng.get_state(), eng.get_degrees()))
    st = eng.stats()
    assert not st["overflow"]
    launches = eng.launch_count()
    eng.close()
    return out, launches, st


@pytest.mark.parametrize("n,episodes,k,hidden", [(100, 1, 3, 32), (100, 1, 1, 32), (64, 2, 2, 16), (128, 1, 4, 64), (25, 5, 3, 32),
                                                 (7, 1, 3, 4)])
def test_mini_path_equals_general_kernels(n, episodes, k, hidden, monkeypatch):
    rng = np.random.default_rng(100 * n + k)
    sd = load_golden("ckpt_n100_k3")["state_dict"] if (k, hidden) == (3, 32) else _random_sd(rng, k, hidden)
    x0 = np.concatenate([flock_env.synthetic_state(n, seed=3 + e, density=1.6) for e in range(episodes)])
    mask = None
    if n == 64:
        mask = np.ones(n * episodes, np.uint8)
        mask[::9] = 0                                    # leaders
    ref, launches_ref, st_ref = _run(monkeypatch, False, n, episodes, k, hidden, sd, x0, mask)
    got, launches, st = _run(monkeypatch, True, n, episodes, k, hidden, sd, x0, mask)
    assert st["step"] == st_ref["step"] and st["n_edges"] == st_ref["n_edges"]
    assert launches < launches_ref                       # one launch per call instead of up to eight per step
    for i, (ra, rb) in enumerate(zip(ref, got)):
        for u, v in zip(ra, rb):
            if episodes > 1 and u.dtype == np.float64 and u.shape[-1:] == (episodes,):
                # per-episode rewards of BATCHED episodes are summed with slotted atomics (arrival order) in both paths
                np.testing.assert_allclose(u, v, rtol=1e-12, atol=1e-15, err_msg=f"record {i}")
            else:
                np.testing.assert_array_equal(u, v, err_msg=f"record {i}")


def test_handles_with_different_stage_depths_share_the_kernels():
    """The dynamic-shared-memory limit is a process-wide attribute of a kernel: a second handle with a shallower neighbour
    stage (or fewer layers) must not lower it under the first handle's feet (found by the compat rollout: its Actor keeps a
    second engine for the dense forward)."""
    from multiagent_gnn_policies_b200.engine import FlockEngine
    sd = load_golden("ckpt_n100_k3")["state_dict"]
    x0 = flock_env.synthetic_state(100, seed=3, density=1.6)
    deep = FlockEngine(n_agents=100, k=3, hidden=32, n_layers=2, comm_radius=1.0, dt=0.01, edge_capacity=99)
    deep.load_state_dict(sd)
    deep.reset(x0)
    shallow = FlockEngine(n_agents=100, k=3, hidden=32, n_layers=2, comm_radius=1.0, dt=0.01, edge_capacity=8)
    shallow.load_state_dict(sd)
    shallow.reset(x0)
    a = np.empty((100, 2), np.float32)
    for eng in (deep, shallow, deep):
        eng.policy(out=a)
        eng.env_step(a)
        eng.rollout(2)
        eng.step(a, None)
    np.testing.assert_array_equal(deep.get_degrees(), deep.get_degrees())
    assert not deep.stats()["overflow"] and not shallow.stats()["overflow"]
    deep.close()
    shallow.close()
