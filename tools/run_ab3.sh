# developer tool: stage times of the default build and of the A/B variants in $VARIANTS (two passes, to see the run-to-run spread)
for pass in 1 2; do
for v in default $VARIANTS; do
  if [ $v = default ]; then unset B200GS_LIB; else export B200GS_LIB=$PWD/variants/libb200gs_$v.so; fi
  timeout 300 python tools/stage_times.py 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); s=d['stages_us']
print('%-12s step %.1f  fwd %.1f  bwd %.1f' % ('$v', d['step_ms']*1000, s['blend_fwd'], s['blend_bwd']))"
done; done
