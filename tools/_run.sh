python tools/gpu_time_applyK.py 160 2>&1 | tail -5
OCMPS_STEP_TRACE=0 python tools/gpu_time_applyK.py 60 2>&1 | tail -3
