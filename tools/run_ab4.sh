# developer tool: like run_ab3.sh on the large workloads (one pass)
for w in stress_train mip360_render; do
for v in default $VARIANTS; do
  if [ $v = default ]; then unset B200GS_LIB; else export B200GS_LIB=$PWD/variants/libb200gs_$v.so; fi
  timeout 600 python tools/stage_times.py --workload $w --steps 5 --views 2 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); s=d['stages_us']
print('%-14s %-10s step %.1f  fwd %.1f  bwd %.1f' % ('$w', '$v', d['step_ms']*1000, s['blend_fwd'], s['blend_bwd']))"
done; done
