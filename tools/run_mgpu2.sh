N=${1:-2}
for nb in 16 64 128; do
B200GS_AR_BLOCKS=$nb timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 2954$((nb%7)) tools/check_exchange.py 2>&1 | grep -E "EXCHANGE|Error|error|Traceback" | head -20
done
