"""Profiling workload at a chosen point of the real cfg2 forward sweep: K steps on bench.py's control, then P steps
between cudaProfilerStart/Stop (run under `ncu --profile-from-start off`).  Step graphs are disabled for the
profiled steps by OCMPS_GRAPH=0 in the environment if plain launches are wanted."""
import sys, os, time, ctypes
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import bench
import optimalcontrolmps_b200 as oc
from optimalcontrolmps_b200.states import ground_state

CFG = bench.CFG
K = int(sys.argv[1]) if len(sys.argv) > 1 else 30
P = int(sys.argv[2]) if len(sys.argv) > 2 else 1
L, d = CFG["L"], CFG["d"]
st = oc.BH_tDMRG(oc.BoseHubbard(L, d), CFG["J"], CFG["tstep"], oc.Args("Cutoff=", CFG["cutoff"], "Maxm=", CFG["maxm"]))
basis, c, u = bench.make_problem_host(0)
pd = st.to_device(ground_state(L, d, CFG["Npart"], CFG["U_i"]))
for k in range(K):
    st.step(pd, u[k], u[k + 1], True)
rt = ctypes.CDLL("/usr/local/cuda/lib64/libcudart.so")
lib = st.ctx.lib
_d = (ctypes.c_ulonglong * 8)()
lib.ocmps_debug_jacobi(_d, 1)
rt.cudaProfilerStart()
t0 = time.perf_counter()
for k in range(K, K + P):
    st.step(pd, u[k], u[k + 1], True)
t1 = time.perf_counter()
rt.cudaProfilerStop()
print("steps", K, "..", K + P, "ms per step", (t1 - t0) / P * 1e3, "dims", pd.bond_dims())
dbg = (ctypes.c_ulonglong * 8)()
lib.ocmps_debug_jacobi(dbg, 0)
nb = max(dbg[4], 1)
print("block SVDs", dbg[1], "sweeps", dbg[0], "max sweeps", dbg[2], "| blocks with >= 64 vectors:", dbg[4], "sweeps/blk", dbg[3] / nb,
      "QR kclk/blk", dbg[5] / nb / 1e3, "Jacobi kclk/blk", dbg[6] / nb / 1e3, "total kclk/blk", dbg[7] / nb / 1e3)
