timeout 300 python tools/stage_times.py 2>gpurun_out/r2_st.err | tee gpurun_out/r2_st_default.json
for v in f7 f8; do
  B200GS_LIB=$PWD/variants/libb200gs_$v.so timeout 300 python tools/stage_times.py 2>/dev/null | tee gpurun_out/r2_st_$v.json
  B200GS_LIB=$PWD/variants/libb200gs_$v.so timeout 600 python tools/stage_times.py --workload stress_train --steps 5 --views 2 2>/dev/null | tee gpurun_out/r2_st_stress_$v.json
done
timeout 600 python tools/profile_e2e.py 2>&1 | grep -E "ms/step"
