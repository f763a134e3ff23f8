# 2-GPU sweep of the custom all-reduce grid size: stand-alone time and the bench step with it
for nb in 16 32 64 256; do
  echo "== B200GS_AR_BLOCKS=$nb"
  B200GS_AR_BLOCKS=$nb timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 2951$((nb % 7)) tools/check_allreduce.py 2>&1 | grep "^P=.*mode=p2p"
  B200GS_ALLREDUCE=p2p B200GS_AR_BLOCKS=$nb timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 2952$((nb % 7)) bench.py --gpus 2 --steps 30 --warmup 5 --no-train 2>/dev/null | python -c "
import json,sys
for l in sys.stdin:
    if l.startswith('{'):
        d=json.loads(l); print('bench ms_per_step', round(d['ms_per_step'],4), 'e2e', round(d['e2e']['ms_per_step'],4))
"
done
