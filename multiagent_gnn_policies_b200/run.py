"""Run one of the reference's own scripts (train.py, test_model.py, ...) UNCHANGED on this engine:

    python -m multiagent_gnn_policies_b200.run /path/to/reference/train.py cfg/dagger.cfg

The compat packages (``gym``, ``gym_flock``, ``learner``) are put first on sys.path, the script's own
directory is appended for its relative paths (cfg/, models/), then the script is executed as __main__.
"""
import os
import runpy
import sys


def main(argv=None):
    argv = list(sys.argv[1:] if argv is None else argv)
    if not argv:
        sys.exit(__doc__)
    script = os.path.abspath(argv[0])
    from multiagent_gnn_policies_b200 import compat
    compat.install()
    sys.argv = [script] + argv[1:]
    os.chdir(os.path.dirname(script))
    # the script's directory must NOT shadow the compat ``learner`` package: append, do not prepend
    if os.path.dirname(script) in sys.path:
        sys.path.remove(os.path.dirname(script))
    sys.path.append(os.path.dirname(script))
    code = compile(open(script).read(), script, "exec")
    exec(code, {"__name__": "__main__", "__file__": script})


if __name__ == "__main__":
    main()
