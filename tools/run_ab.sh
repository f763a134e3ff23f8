timeout 300 python tools/stage_times.py 2>gpurun_out/r2_st.err | tee gpurun_out/r2_st_default.json
timeout 600 python tools/stage_times.py --workload stress_train --steps 5 --views 2 2>/dev/null | tee gpurun_out/r2_st_stress.json
timeout 600 python tools/stage_times.py --workload mip360_render --steps 5 --views 2 --fwd-only 2>/dev/null | tee gpurun_out/r2_st_mip.json
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -3
