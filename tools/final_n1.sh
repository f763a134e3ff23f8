set -x
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -3
timeout 300 python __graft_entry__.py smoke 2>&1 | tail -2
timeout 900 python bench.py --impl reference > gpurun_out/r02b_final_bench_reference_n1.json 2> gpurun_out/r02b_final_bench_reference_n1.err; tail -2 gpurun_out/r02b_final_bench_reference_n1.err
timeout 900 python bench.py > gpurun_out/r02b_final_bench_n1.json 2> gpurun_out/r02b_final_bench_n1.err; tail -2 gpurun_out/r02b_final_bench_n1.err
python - <<'PY'
import json
d=json.loads(open('gpurun_out/r02b_final_bench_n1.json').read().strip().splitlines()[-1])
r=json.loads(open('gpurun_out/r02b_final_bench_reference_n1.json').read().strip().splitlines()[-1])
print("ours ms", d["ms_per_step"], "e2e", d["e2e"]["ms_per_step"], "vanilla", d["vanilla"]["ms_per_view"], d["vanilla"]["e2e"]["ms_per_step"], "train", d["train"]["ms_per_iter"])
print("ref  ms", r["ms_per_step"], "e2e", r["e2e"]["ms_per_step"], "vanilla", r["vanilla"]["ms_per_view"], r["vanilla"]["e2e"]["ms_per_step"], "train", r["train"]["ms_per_iter"])
print("stages", {k: round(v*1000,1) for k,v in d["roofline"]["stage_ms"].items()}, "roofline", d["roofline"]["kernel"], round(d["roofline"]["frac"],4), d["roofline"]["issue"] and round(d["roofline"]["issue"]["frac"],3), "step frac", round(d["roofline"]["step"]["frac"],3))
print("render_sharded", d.get("render_sharded",{}).get("views_per_s"), "stress", d.get("stress_train",{}).get("ms_per_step"), "cpu", d["cpu_baseline"], "clocks", d["clocks"], "launches/step", d["gpu_launches_per_step"])
PY
CMD="python bench.py --steps 3 --warmup 3 --no-extras --no-cpu-baseline"
timeout 300 $CMD > gpurun_out/r02b_ncu_plain.log 2>&1 && \
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 900 --csv --log-file gpurun_out/r02b_final_launches_raw.csv $CMD > gpurun_out/r02b_ncu_launches.log 2>&1
tail -2 gpurun_out/r02b_ncu_launches.log
grep -c libb200gs gpurun_out/r02b_final_bench_reference_n1.err
