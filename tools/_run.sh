OCMPS_STEP_TRACE=1 python tools/gpu_prof_at.py 175 3 2>&1 | grep -v "^block" | tail -5
OCMPS_STEP_TRACE=1 python tools/gpu_prof_at.py 30 2 2>&1 | grep "ocmps step" | tail -2
