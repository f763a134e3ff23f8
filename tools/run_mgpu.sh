N=${1:-2}
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29541 tools/check_exchange.py 2>&1 | grep -E "EXCHANGE|Error|error|Traceback" | head -5 | tee gpurun_out/r02_exchange_n$N.txt
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29543 bench.py --gpus $N > gpurun_out/r02_bench_n$N.json 2> gpurun_out/r02_bench_n$N.err; tail -2 gpurun_out/r02_bench_n$N.err
python - <<PY
import json
d=json.loads(open('gpurun_out/r02_bench_n$N.json').read().strip().splitlines()[-1])
print("N=$N value", round(d["value"],1), "ms", round(d["ms_per_step"],4), "e2e", round(d["e2e"]["value"],1), "train it/s", round(d["train"]["iters_per_s"],1), "render_sharded", round(d["render_sharded"]["views_per_s"],1), "stress ms/step", round(d["stress_train"]["ms_per_step"],2), "check", d["collective_check"]["max_rel_err"], d["collective_check"]["identical_on_all_ranks"])
PY
