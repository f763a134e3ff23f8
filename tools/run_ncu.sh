CMD="python tools/stage_times.py --no-graph --steps 3"
timeout 300 $CMD > gpurun_out/r2_ncu_plain.log 2>&1 && \
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"preprocess|onesweep|scan_emit" -s 27 -c 9 -o gpurun_out/r2_front_v1 -f $CMD > gpurun_out/r2_ncu.log 2>&1
tail -3 gpurun_out/r2_ncu.log
