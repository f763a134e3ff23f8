# developer tool: parity spot checks, stage times of the default build and of A/B variants, then the GPU test suite
cd tests && timeout 300 python gpu_check.py tiny_sh3_ext small_sh3 inside_sh2_white_ext 2>&1 | grep -E "==|product vs oracle" -A2 | grep -E "==|int-mismatch|grad rel" | head; cd ..
timeout 300 python tools/stage_times.py 2>gpurun_out/st.err | tee gpurun_out/r2b_st_new.json
tail -2 gpurun_out/st.err
for v in $VARIANTS; do
  B200GS_LIB=$PWD/variants/libb200gs_$v.so timeout 300 python tools/stage_times.py 2>/dev/null | tee gpurun_out/r2b_st_$v.json
done
timeout 600 python tools/stage_times.py --workload stress_train --steps 5 --views 2 2>/dev/null | tee gpurun_out/r2b_st_stress_new.json
timeout 1200 python -m pytest tests -m gpu -x -q 2>&1 | tail -6
