timeout 900 python bench.py --steps 30 --warmup 5 > gpurun_out/r2_bench_a.json 2> gpurun_out/r2_bench_a.err; tail -5 gpurun_out/r2_bench_a.err
python - <<'PY'
import json
d=json.loads(open('gpurun_out/r2_bench_a.json').read().strip().splitlines()[-1])
for k in ("ms_per_step","value","vanilla","train","render_sharded","stress_train","collective_check"): print(k, d.get(k))
print("e2e", d["e2e"]["ms_per_step"], "stages", {k: round(v*1000,1) for k,v in d["roofline"]["stage_ms"].items()})
PY
timeout 600 python bench.py --impl reference --steps 30 --warmup 5 > gpurun_out/r2_bench_ref_a.json 2> gpurun_out/r2_bench_ref_a.err; tail -3 gpurun_out/r2_bench_ref_a.err
python - <<'PY'
import json
d=json.loads(open('gpurun_out/r2_bench_ref_a.json').read().strip().splitlines()[-1])
for k in ("ms_per_step","value","vanilla","train"): print("REF", k, d.get(k))
print("REF e2e", d["e2e"])
PY
grep -c libb200gs /proc/self/maps
