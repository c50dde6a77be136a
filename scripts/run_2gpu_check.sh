timeout 900 python -m pytest tests/test_gpu_sharding_ranks.py -q 2>&1 | tail -4
for fold in 1 0; do echo "== FGNN_SHARD_FOLD=$fold"; FGNN_SHARD_FOLD=$fold timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus 2 --steps 200 --warmup 20 2>&1 | grep -E "^\{|\[profile\]" | python -c "
import sys,json
for l in sys.stdin:
    if l.startswith('{'):
        l=json.loads(l); print({k:l.get(k) for k in ('value','ms_per_step','halo_records_per_step')}, 'e2e', l['e2e']['value'], 'strong', json.dumps(l.get('strong'))[:600])
    else: print(l.strip()[:400])
"; done
