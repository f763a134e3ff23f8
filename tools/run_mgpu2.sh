N=${1:-2}
for tma in 1; do
echo "SCATTER_TMA=$tma"
B200GS_SCATTER_TMA=$tma timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 2954$tma tools/check_exchange.py 2>&1 | grep -E "EXCHANGE|Error|error|Traceback" | head -20
done
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29547 tools/check_allreduce.py 2>&1 | grep -E "^P=|Error|Traceback" | head -4
