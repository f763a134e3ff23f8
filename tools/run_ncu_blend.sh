# developer tool: ncu --set full of the two blend kernels (one launch each), optionally at a forced CTAs/SM ($1 = "f,b")
TAG=${2:-default}
CMD="python tools/stage_times.py --no-graph --steps 3"
[ -n "$1" ] && export B200GS_BLEND_CTAS=$1
timeout 300 $CMD > gpurun_out/ncu_plain_$TAG.log 2>&1 && \
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"blend_" -s 6 -c 2 -o gpurun_out/blend_$TAG -f $CMD > gpurun_out/ncu_$TAG.log 2>&1
tail -2 gpurun_out/ncu_$TAG.log
