for lag in 0 1 2; do
B200GS_E2E_LAG=$lag timeout 600 python bench.py --steps 60 --warmup 5 --no-extras --no-cpu-baseline --no-train 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('lag $lag: ms_per_step', round(d['ms_per_step'],4), 'e2e', round(d['e2e']['ms_per_step'],4), 'vanilla e2e', round(d['vanilla']['e2e']['ms_per_step'],4))"
done
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -3
