# developer tool: stage times of the default build under several B200GS_BLEND_CTAS settings ($CTAS_LIST, "f,b" each)
for pass in 1 2; do
for c in $CTAS_LIST; do
  B200GS_BLEND_CTAS=$c timeout 300 python tools/stage_times.py 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); s=d['stages_us']
print('ctas %-6s step %.1f  fwd %.1f  bwd %.1f' % ('$c', d['step_ms']*1000, s['blend_fwd'], s['blend_bwd']))"
done; done
