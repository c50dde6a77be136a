"""Recipe: stage the UNMODIFIED reference learner under oracle/_ref/ so that it travels to the GPU box.

TEST / BENCH INFRASTRUCTURE ONLY.  `/root/reference` exists in the build container but not on the GPU
box, and `bench.py --impl reference` / `cpu_baseline` must time the reference's own CPU implementation of
the path (learner/state_with_delay.py:6-53 MultiAgentStateWithDelay + learner/gnn_dagger.py:55-72
DAGGER.select_action + learner/actor.py:45-86 Actor.forward).  This script copies the files of the
reference's `learner` package, byte for byte, from where they lie into `oracle/_ref/learner/`
(`oracle/_ref/` is git-ignored -- nothing of the reference enters the history -- but not gpurun-ignored,
like a built `.so`), and records their sha256 in `oracle/_ref/MANIFEST.json`.

    python -m oracle.make_ref            # run by __graft_entry__.build() when /root/reference is present

The product never imports oracle/_ref: only bench.py's reference leg and tests/ do.
"""
import hashlib
import json
import os
import shutil
import sys

REF = os.environ.get("FGNN_REFERENCE", "/root/reference")
HERE = os.path.dirname(os.path.abspath(__file__))
DST = os.path.join(HERE, "_ref")
FILES = ["learner/__init__.py", "learner/actor.py", "learner/state_with_delay.py", "learner/gnn_dagger.py",
         "learner/replay_buffer.py", "cfg/dagger.cfg",
         # the reference's own entry scripts and shipped checkpoint: tests/test_gpu_reference_scripts.py runs them UNCHANGED
         # through multiagent_gnn_policies_b200.run (the learner modules they import are then the compat ones)
         "test_model.py", "train.py", "models/actor_FlockingRelative-v0_dagger_k3"]


def make(ref=REF, dst=DST, quiet=False):
    """Returns True when oracle/_ref holds the reference files (copied now or already there)."""
    if not os.path.isdir(os.path.join(ref, "learner")):
        return os.path.exists(os.path.join(dst, "MANIFEST.json"))
    manifest = {"source": ref, "files": {}}
    for rel in FILES:
        src = os.path.join(ref, rel)
        out = os.path.join(dst, rel)
        os.makedirs(os.path.dirname(out), exist_ok=True)
        shutil.copyfile(src, out)
        manifest["files"][rel] = hashlib.sha256(open(out, "rb").read()).hexdigest()
    json.dump(manifest, open(os.path.join(dst, "MANIFEST.json"), "w"), indent=1)
    if not quiet:
        print(f"[oracle/_ref] staged {len(FILES)} reference files from {ref}")
    return True


def available(dst=DST):
    return os.path.exists(os.path.join(dst, "learner", "gnn_dagger.py"))


def verify(dst=DST):
    """sha256 of every staged file against the manifest written at copy time."""
    m = json.load(open(os.path.join(dst, "MANIFEST.json")))
    return all(hashlib.sha256(open(os.path.join(dst, rel), "rb").read()).hexdigest() == h for rel, h in m["files"].items())


if __name__ == "__main__":
    sys.exit(0 if make() else 1)
