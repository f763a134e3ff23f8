cd tests && timeout 300 python gpu_check.py tiny_sh3_ext small_sh3 inside_sh2_white_ext 2>&1 | grep -E "==|product vs oracle" -A2 | grep -E "==|int-mismatch|grad rel" | head; cd ..
timeout 300 python tools/stage_times.py 2>gpurun_out/r2_st.err | tee gpurun_out/r2_st_default.json
tail -2 gpurun_out/r2_st.err
timeout 600 python tools/stage_times.py --workload dtu_scan_3view --steps 10 2>/dev/null | tee gpurun_out/r2_st_dtu.json
timeout 600 python tools/stage_times.py --workload stress_train --steps 5 --views 2 2>/dev/null | tee gpurun_out/r2_st_stress.json
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -4
