timeout 900 python -m pytest tests/test_gpu_sharding_ranks.py -q 2>&1 | tail -6
for halo in p2p gather; do echo "== halo $halo"; FGNN_HALO=$halo timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus 2 --steps 200 --warmup 20 2>&1 | tail -1 | python -c "
import sys,json
l=json.loads(sys.stdin.readline()); print({k:l[k] for k in ('value','ms_per_step','halo_records_per_step','strong')}, l['e2e']['value'])"; done
