"""K|psi> (exactApplyMPO, 2*chi intermediate bonds) on a saturated cfg2 state: wall time per call (GPU)."""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import bench
import optimalcontrolmps_b200 as oc
from optimalcontrolmps_b200.states import ground_state

CFG = bench.CFG
K = int(sys.argv[1]) if len(sys.argv) > 1 else 160
L, d = CFG["L"], CFG["d"]
st = oc.BH_tDMRG(oc.BoseHubbard(L, d), CFG["J"], CFG["tstep"], oc.Args("Cutoff=", CFG["cutoff"], "Maxm=", CFG["maxm"]))
basis, c, u = bench.make_problem_host(0)
pd = st.to_device(ground_state(L, d, CFG["Npart"], CFG["U_i"]))
for k in range(K):
    st.step(pd, u[k], u[k + 1], True)
for rep in range(3):
    t0 = time.perf_counter()
    kp = st.exactApplyMPO(pd)
    print("apply_K ms", (time.perf_counter() - t0) * 1e3, "dims", max(kp.bond_dims()), flush=True)
t0 = time.perf_counter()
for k in range(K, K + 5):
    st.step(pd, u[k], u[k + 1], True)
print("step ms", (time.perf_counter() - t0) / 5 * 1e3)
