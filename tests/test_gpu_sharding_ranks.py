"""Real ranks: one process per GPU (torchrun), every halo transport, against a single unsharded engine bit for bit.
Runs for every world size in {2, 4, 8} that the box has devices for; with one GPU visible it is skipped (the in-process
version of the same protocol is tests/test_gpu_sharding.py)."""
import os
import subprocess
import sys

import pytest

from conftest import ROOT

pytestmark = pytest.mark.gpu


def _device_count():
    import torch
    return torch.cuda.device_count() if torch.cuda.is_available() else 0


@pytest.mark.parametrize("world", [2, 4, 8])
@pytest.mark.parametrize("mode", ["p2p", "p2p_host", "gather"])
def test_real_ranks_reproduce_single_engine(world, mode):
    if _device_count() < world:
        pytest.skip(f"needs {world} GPUs")
    port = 29600 + world + {"p2p": 10, "p2p_host": 20, "gather": 0}[mode]
    env = dict(os.environ, FGNN_CHECK_N="120000", FGNN_CHECK_STEPS="40")
    if mode != "gather":
        env["FGNN_STEP_MODE"] = "1"      # the warp-tiled adjacency on sharded pools too (by default it starts at larger shards)
    out = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={world}",
                          "--master-addr", "127.0.0.1", "--master-port", str(port),
                          os.path.join(ROOT, "scripts", "check_sharded_nccl.py"), mode],
                         capture_output=True, text=True, timeout=900, env=env, cwd=ROOT)
    assert out.returncode == 0, (out.stdout[-2000:], out.stderr[-3000:])
    assert f"sharded rollout ({mode}) == single engine: True" in out.stdout, out.stdout[-2000:]
    assert "one owner per agent: True" in out.stdout and "overflow 0" in out.stdout
