#!/bin/bash
# Build a -DOCMPS_JAC_TRACE variant of the library into a side path, run one traced step of the real cfg2 sweep with it on the
# GPU box (per-block shapes, ranks and cycle counts of the QR and Jacobi phases), restore the normal library.
# usage: tools/trace_build_run.sh [steps_before] [extra env, e.g. OCMPS_CLUSTER_JACOBI=1]
set -e
cd /root/repo
K=${1:-175}
ENVX=${2:-}
mkdir -p build/trace
nvcc -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -std=c++17 -Xcompiler -fPIC -DOCMPS_JAC_TRACE -c optimalcontrolmps_b200/csrc/decomp.cu -o build/trace/decomp.o
cp optimalcontrolmps_b200/libocmps.so build/trace/libocmps_good.so
nvcc -gencode arch=compute_100a,code=sm_100a -shared -o optimalcontrolmps_b200/libocmps.so optimalcontrolmps_b200/csrc/zgemm.o build/trace/decomp.o optimalcontrolmps_b200/csrc/elementwise.o optimalcontrolmps_b200/csrc/engine.o -lcudart
/usr/local/graft/bin/gpurun --timeout 900 -- "OCMPS_GRAPH=0 $ENVX timeout 600 python ${TOOL:-tools/gpu_prof_at.py} $K 1 > gpurun_out/jt.log 2>&1; tail -2 gpurun_out/jt.log" 2>&1 | tail -4
cp build/trace/libocmps_good.so optimalcontrolmps_b200/libocmps.so
