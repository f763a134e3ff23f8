set -x
CMD="python tools/stage_times.py --no-graph --steps 3"
timeout 300 $CMD > gpurun_out/r2_ncu_plain.log 2>&1 && \
timeout 900 ncu --set full --clock-control none --import-source on -k regex:blend_ -s 7 -c 2 -o gpurun_out/r2_blend_v2 -f $CMD > gpurun_out/r2_ncu.log 2>&1
tail -5 gpurun_out/r2_ncu.log
