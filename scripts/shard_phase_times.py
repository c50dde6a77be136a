"""torchrun helper: time the three phases of a sharded step (begin graph / all-gather / end graph) separately."""
import os, sys
import numpy as np, torch, torch.distributed as dist
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from bench import make_workload, make_weights, DENSITY
from multiagent_gnn_policies_b200 import parallel

rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
local = int(os.environ.get("LOCAL_RANK", rank))
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
N, K, R = 1_000_000, 3, 1.0
side = np.sqrt(N / DENSITY)
n_total = world * N
ranges = parallel.shard_ranges(n_total, world)
lo, cnt = ranges[rank]
x = np.zeros((n_total, 4)); x[:, 0] = parallel.FAR
for q in (rank - 1, rank, rank + 1):
    if 0 <= q < world:
        x[ranges[q][0]:ranges[q][0] + ranges[q][1]] = make_workload(N, seed=11 + q, x_offset=q * side, bias_seed=11)
depth = parallel.halo_depth(K, R)
cap = int(2.0 * (depth + 2 * R) * side * DENSITY * 2) + 1024
sd, _ = make_weights(32, 3, 2)
be = parallel.CudaShardBackend(n_total, lo, cnt, ghost_capacity=2 * cap, device=local, k=K, hidden=32, n_layers=2,
                               comm_radius=R, dt=0.01, edge_capacity=32, grid_dim=int(np.ceil((side + 2 * depth + 4))) + 2,
                               grid_dim_y=int(np.ceil(side)) + 4)
be.engine.load_state_dict(sd)
flock = parallel.ShardedFlock(be, rank, world, K, R, cap, parallel.nccl_all_gather(world, cap, be.device))
bounds = np.array([-parallel.INF] + [q * side for q in range(1, world)] + [parallel.INF])
flock.reset(x, ranges, bounds=bounds, frame_velocity=float(np.random.default_rng(11).uniform(-3, 3, size=(2,))[0]))
for _ in range(20):
    flock.step()
stride = (flock.cap + 1) * parallel.RECORD
ev = [torch.cuda.Event(enable_timing=True) for _ in range(4)]
acc = np.zeros(3)
n = 50
for _ in range(n):
    dist.barrier(); torch.cuda.synchronize()
    ev[0].record()
    be.step_begin(flock.recv.reshape(-1)[1:], stride, flock.send, flock.cap)
    ev[1].record()
    flock.recv = flock.all_gather(flock.send)
    ev[2].record()
    be.step_end(flock.recv, flock.cap)
    ev[3].record()
    torch.cuda.synchronize()
    acc += [ev[0].elapsed_time(ev[1]), ev[1].elapsed_time(ev[2]), ev[2].elapsed_time(ev[3])]
if rank == 0:
    print("world", world, "ms per step: begin(hops+final+pack) %.4f  all-gather %.4f  end(unpack+build) %.4f  sum %.4f" % (*(acc / n), acc.sum() / n),
          "| records/rank", int(flock.recv[rank, 0, 0].item()), "bytes/rank", (flock.cap + 1) * parallel.RECORD * 8)
dist.destroy_process_group()
