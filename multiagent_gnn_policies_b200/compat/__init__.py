"""Drop-in call surface of the reference: importable packages named ``gym``, ``gym_flock`` and
``learner`` whose hot path runs on the CUDA engine.

    import multiagent_gnn_policies_b200.compat as compat
    compat.install()          # puts this directory first on sys.path
    import gym, gym_flock     # -> the shims below
    from learner.gnn_dagger import DAGGER, train_dagger

``python -m multiagent_gnn_policies_b200.run <reference>/train.py cfg/dagger.cfg`` does the same and
then executes the reference's own script unchanged (train.py:45-63, test_model.py:50-67).
"""
import os
import sys

COMPAT_DIR = os.path.dirname(os.path.abspath(__file__))


def install():
    if COMPAT_DIR not in sys.path:
        sys.path.insert(0, COMPAT_DIR)
    for name in ("gym", "gym_flock", "learner"):
        mod = sys.modules.get(name)
        if mod is not None and not getattr(mod, "__file__", "").startswith(COMPAT_DIR):
            for key in [k for k in sys.modules if k == name or k.startswith(name + ".")]:
                del sys.modules[key]
