"""Profiling workload for ncu (`--profile-from-start off`): ONE cost+gradient evaluation over a short horizon (Nt points) that starts
from the psi of bench.py's cfg2 control after K0 Trotter steps (bond dimensions saturated at 100) and ends at the U=50 ground state:
forward sweep + slice store, backward sweep, batched transfer-matrix overlaps (divT with the K MPO, fidelities).
usage: gpu_prof_eval.py [K0=170] [Nt=8]"""
import sys, os, time, ctypes
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import bench
import optimalcontrolmps_b200 as oc
from optimalcontrolmps_b200.states import ground_state

CFG = bench.CFG
K0 = int(sys.argv[1]) if len(sys.argv) > 1 else 170
Nt = int(sys.argv[2]) if len(sys.argv) > 2 else 8
L, d = CFG["L"], CFG["d"]
st = oc.BH_tDMRG(oc.BoseHubbard(L, d), CFG["J"], CFG["tstep"], oc.Args("Cutoff=", CFG["cutoff"], "Maxm=", CFG["maxm"]))
basis, c, u = bench.make_problem_host(0)
start = st.to_device(ground_state(L, d, CFG["Npart"], CFG["U_i"]))
for k in range(K0):
    st.step(start, u[k], u[k + 1], True)
# a target with real entanglement too: the same state evolved a little further (so that xi is as heavy as psi)
target = st.new_mps().copy_from(start)
for k in range(K0, K0 + 3):
    st.step(target, u[k], u[k + 1], True)
uw = list(u[K0:K0 + Nt])
p = oc.OptimalControl(target, start, st, Nt, CFG["gamma"])
p.setThreadCount(2)
p.getAnalyticGradient(uw, True)          # warm: workspaces, graphs
rt = ctypes.CDLL("/usr/local/cuda/lib64/libcudart.so")
rt.cudaProfilerStart()
t0 = time.perf_counter()
g = p.getAnalyticGradient(uw, True)
cost = p.getCost(uw, False)
t1 = time.perf_counter()
rt.cudaProfilerStop()
print("eval over", Nt, "points from step", K0, ":", (t1 - t0) * 1e3, "ms, cost", cost, "max dims", int(p.psi_t.bond_dims().max()), int(p.xi_t.bond_dims().max()))
