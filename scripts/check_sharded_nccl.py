"""torchrun check of the multi-GPU halo paths on REAL ranks (run on >= 2 GPUs; tests/test_gpu_sharding_ranks.py spawns it):
    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 \
        scripts/check_sharded_nccl.py [p2p|p2p_host|gather|native]
Every rank shards one flock by index, rolls out T steps, and rank 0 compares the owned slices of all ranks with a single
unsharded engine, bit for bit.  Transports: p2p (default: records stored straight into the peers' inboxes over NVLink, one
CUDA graph per step), gather (torch.distributed all-gather between two graph halves), native (ncclAllGather inside the graph)."""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from multiagent_gnn_policies_b200 import parallel                 # noqa: E402
from multiagent_gnn_policies_b200.engine import FlockEngine       # noqa: E402
from bench import make_workload                                   # noqa: E402


def main():
    rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
    local = int(os.environ.get("LOCAL_RANK", rank))
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    mode = sys.argv[1] if len(sys.argv) > 1 else "p2p"
    n_total, steps, K, R = int(os.environ.get("FGNN_CHECK_N", "200000")), int(os.environ.get("FGNN_CHECK_STEPS", "60")), 3, 1.0
    g = np.load(os.path.join(ROOT, "tests", "golden", "ckpt_n100_k3.npz"))
    sd = {k[3:]: g[k] for k in g.files if k.startswith("sd.")}
    x0 = make_workload(n_total, seed=5)
    x0 = x0[np.argsort(x0[:, 0], kind="stable")]
    ranges = parallel.shard_ranges(n_total, world)
    lo, cnt = ranges[rank]
    cap = 20000
    be = parallel.CudaShardBackend(n_total, lo, cnt, ghost_capacity=4 * cap, device=local, k=K, hidden=32, n_layers=2,
                                   comm_radius=R, dt=0.01, edge_capacity=48)
    be.engine.load_state_dict(sd)
    flock = parallel.ShardedFlock(be, rank, world, K, R, cap, parallel.nccl_all_gather(world, cap, be.device))
    flock.reset(x0, ranges)
    if mode == "native":
        be.init_comm(rank, world)              # one CUDA graph per step with ncclAllGather inside
    elif mode in ("p2p", "p2p_host"):
        flock.enable_p2p(parallel.torch_all_gather_object(world))
    if mode == "p2p_host":
        # the reference-facing loop on every rank: select_action -> pinned host array -> env.step(host array), halo over p2p
        # (bit-identical to the closed loop: the API-split path leaves the same bits as the fused step)
        act_host = torch.empty((be.engine.rows_io, 2), dtype=torch.float32, pin_memory=True).numpy()
        for _ in range(steps):
            flock.step_host(act_host)
    else:
        for _ in range(steps):
            flock.step()
    torch.cuda.synchronize()
    if os.environ.get("FGNN_CHECK_PROFILE") == "1" and mode == "p2p":
        # the p2p step kernel by kernel (CUDA events, un-graphed) on every rank, then the graph-replayed step time
        per = {}
        for _ in range(10):
            for name, ms in be.engine.profile_step():
                per[name] = per.get(name, 0.0) + ms / 10
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        dist.barrier()
        torch.cuda.synchronize()
        e0.record()
        for _ in range(200):
            flock.step()
        e1.record()
        torch.cuda.synchronize()
        print(f"rank {rank}: graph step {e0.elapsed_time(e1) / 200 * 1e3:.1f} us | per kernel (events): "
              + " ".join(f"{k}={v * 1e3:.1f}" for k, v in per.items()) + f" | ghosts {be.engine.stats()['n_ghosts']}", flush=True)
        steps += 210
    ids, st = be.owned_state()
    # assemble the global state on every rank: sum of each rank's owned rows (exactly one owner per agent)
    x_all = torch.zeros((n_total, 4), dtype=torch.float64, device="cuda")
    cnt_all = torch.zeros((n_total,), dtype=torch.float64, device="cuda")
    x_all[torch.from_numpy(ids).cuda().long()] = torch.from_numpy(st).cuda()
    cnt_all[torch.from_numpy(ids).cuda().long()] = 1.0
    dist.all_reduce(x_all)
    dist.all_reduce(cnt_all)
    ok = True
    if rank == 0:
        single = FlockEngine(n_agents=n_total, k=K, hidden=32, n_layers=2, comm_radius=R, dt=0.01, edge_capacity=48,
                             device=local)
        single.load_state_dict(sd)
        single.reset(x0)
        single.rollout(steps)
        ref = single.get_state()
        one_owner = bool((cnt_all == 1).all().item())
        ok = one_owner and np.array_equal(x_all.cpu().numpy(), ref)
        print(f"sharded rollout ({mode}) == single engine:", ok, "| one owner per agent:", one_owner, "| world", world,
              "| owned now", ids.size, "of initial", cnt, "| overflow", be.overflow())
    dist.barrier()
    dist.destroy_process_group()
    if not ok:
        sys.exit(1)


if __name__ == "__main__":
    main()
