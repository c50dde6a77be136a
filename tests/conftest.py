import glob
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

GOLDEN_DIR = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


def pytest_collection_modifyitems(config, items):
    """`pytest tests` on a box without a GPU (or without the built library) skips the gpu-marked tests instead of failing
    inside them; on the B200 box nothing is skipped (a missing libfgnn.so there is an error, not a skip)."""
    import torch
    if torch.cuda.is_available():
        return
    skip = pytest.mark.skip(reason="needs a CUDA device (no GPU here: run with -m gpu on the B200 box)")
    for item in items:
        if "gpu" in item.keywords:
            item.add_marker(skip)


def _all_golden():
    return sorted(os.path.splitext(os.path.basename(p))[0] for p in glob.glob(os.path.join(GOLDEN_DIR, "*.npz")))


def golden_names():
    """Rollout fixtures (oracle/gen_golden.py)."""
    return [n for n in _all_golden() if not n.startswith(("train_", "agg"))]


def agg_golden_names():
    """Actor.forward fixtures for ind_agg != 0 / unequal layer widths (oracle/gen_golden_agg.py)."""
    return [n for n in _all_golden() if n.startswith("agg")]


def load_agg_golden(name):
    g = dict(np.load(os.path.join(GOLDEN_DIR, name + ".npz")))
    g["state_dict"] = {k[3:]: v for k, v in g.items() if k.startswith("sd.")}
    for key in ("n_agents", "k", "ind_agg", "batch", "seed"):
        g[key] = int(g[key])
    g["layers"] = [int(v) for v in g["layers"]]
    return g


def train_golden_names():
    """gradient_step fixtures (oracle/gen_golden_train.py)."""
    return [n for n in _all_golden() if n.startswith("train_")]


def load_train_golden(name):
    g = dict(np.load(os.path.join(GOLDEN_DIR, name + ".npz")))
    for tag in ("sd0", "sd1", "sdT", "grad1"):
        g[tag] = {k[len(tag) + 1:]: v for k, v in g.items() if k.startswith(tag + ".")}
    for key in ("n_agents", "k", "hidden", "n_layers", "batch", "steps", "seed"):
        g[key] = int(g[key])
    g["lr"] = float(g["lr"])
    return g


def load_golden(name):
    g = dict(np.load(os.path.join(GOLDEN_DIR, name + ".npz")))
    sd = {k[3:]: v for k, v in g.items() if k.startswith("sd.")}
    g["state_dict"] = sd
    for key in ("n_agents", "k", "hidden", "n_layers", "steps", "seed"):
        g[key] = int(g[key])
    for key in ("comm_radius", "dt", "v_max"):
        g[key] = float(g[key])
    return g


@pytest.fixture(params=golden_names())
def golden(request):
    return load_golden(request.param)


@pytest.fixture(params=train_golden_names())
def train_golden(request):
    return load_train_golden(request.param)


def rel_inf(a, b):
    """||a-b||_inf / ||b||_inf, the parity metric of SURVEY.md section 7.3 item 4."""
    a = np.asarray(a, dtype=np.float64)
    b = np.asarray(b, dtype=np.float64)
    den = np.abs(b).max()
    return float(np.abs(a - b).max() / (den if den > 0 else 1.0))
