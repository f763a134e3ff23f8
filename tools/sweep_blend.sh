for cfg in "1 1" "2 2" "4 4" "2 4" "4 2"; do set -- $cfg; echo "== FWD_NW=$1 BWD_NW=$2"; B200GS_FWD_NW=$1 B200GS_BWD_NW=$2 timeout 200 python bench.py --steps 20 --warmup 3 --no-cpu-baseline 2>/dev/null | python -c "
import json,sys
for l in sys.stdin:
    if l.startswith('{'):
        d=json.loads(l); print(round(d['ms_per_step'],4), 'e2e', round(d['e2e']['ms_per_step'],4), {k:round(v,4) for k,v in d['roofline']['stage_ms'].items() if 'blend' in k})
"; done
cd tests; timeout 120 python gpu_check.py tiny_sh3_ext small_sh3 2>&1 | grep -E "product vs" | head -4
