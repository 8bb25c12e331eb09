#!/bin/bash
# Build a -DOCMPS_JAC_TRACE variant of the library into a side path, run the step profiler with it on the GPU box, restore.
set -e
cd /root/repo
mkdir -p build/trace
nvcc -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -std=c++17 -Xcompiler -fPIC -DOCMPS_JAC_TRACE -c optimalcontrolmps_b200/csrc/decomp.cu -o build/trace/decomp.o
cp optimalcontrolmps_b200/libocmps.so build/trace/libocmps_good.so
nvcc -gencode arch=compute_100a,code=sm_100a -shared -o optimalcontrolmps_b200/libocmps.so optimalcontrolmps_b200/csrc/zgemm.o build/trace/decomp.o optimalcontrolmps_b200/csrc/elementwise.o optimalcontrolmps_b200/csrc/engine.o -lcudart
/usr/local/graft/bin/gpurun --timeout 900 -- 'OCMPS_GRAPH=0 timeout 600 python tools/gpu_prof_steps.py 60 1 > gpurun_out/jt.log 2>&1; tail -1 gpurun_out/jt.log' 2>&1 | tail -3
cp build/trace/libocmps_good.so optimalcontrolmps_b200/libocmps.so
