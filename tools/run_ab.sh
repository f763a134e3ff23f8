set -x
cd tests && timeout 300 python gpu_check.py tiny_sh3_ext small_sh3 inside_sh2_white_ext small_precomp 2>&1 | grep -E "==|product vs oracle|product vs reference" -A2 | grep -E "==|int-mismatch|grad rel" | head -40; cd ..
timeout 300 python tools/stage_times.py 2>gpurun_out/r2_st.err | tee gpurun_out/r2_st_default.json
for c in 6,5 6,4 6,3 6,2 4,3 3,3; do
  echo "CTAS=$c"; B200GS_BLEND_CTAS=$c timeout 300 python tools/stage_times.py 2>/dev/null | tee gpurun_out/r2_st_ctas_$c.json
done
timeout 600 python tools/stage_times.py --workload stress_train --steps 5 --views 2 2>/dev/null | tee gpurun_out/r2_st_stress.json
timeout 600 python tools/stage_times.py --workload dtu_scan_3view --steps 10 2>/dev/null | tee gpurun_out/r2_st_dtu.json
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -3
