# A/B of the blend kernel variants on the default bench workload (run on the GPU box)
for cfg in "$@"; do echo "== B200GS_BLEND=$cfg"; B200GS_BLEND=$cfg timeout 200 python bench.py --steps 20 --warmup 3 --no-cpu-baseline 2>/dev/null | python -c "
import json,sys
for l in sys.stdin:
    if l.startswith('{'):
        d=json.loads(l); print(round(d['ms_per_step'],4), 'e2e', round(d['e2e']['ms_per_step'],4), {k:round(v,4) for k,v in d['roofline']['stage_ms'].items() if 'blend' in k})
"; (cd tests; B200GS_BLEND=$cfg timeout 120 python gpu_check.py tiny_sh3_ext small_sh3 inside_sh1 2>&1 | grep -E "product vs" | head -6); done
