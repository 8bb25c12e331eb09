"""Time <N_j>, <N_j(N_j-1)>, <N_j^2> over all resident slices of a cfg2 forward sweep."""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import bench
import optimalcontrolmps_b200 as oc
from optimalcontrolmps_b200.states import ground_state

CFG = bench.CFG
ctx = oc.Context.default(0)
st = oc.BH_tDMRG(oc.BoseHubbard(CFG["L"], CFG["d"]), CFG["J"], CFG["tstep"], oc.Args("Cutoff=", CFG["cutoff"], "Maxm=", CFG["maxm"]), ctx=ctx)
psi_i, psi_f = ground_state(CFG["L"], CFG["d"], CFG["Npart"], CFG["U_i"]), ground_state(CFG["L"], CFG["d"], CFG["Npart"], CFG["U_f"])
basis, c, u = bench.make_problem_host(0)
o = oc.OptimalControl(psi_f, psi_i, st, basis, CFG["gamma"])
o.getCost(list(c), True)
for rep in range(3):
    t0 = time.perf_counter()
    v, nrm = o.psi_t.expectationValues(("N", "N(N-1)", "NN"), return_norm=True)
    dt = time.perf_counter() - t0
    print("slices", v.shape[0], "seconds", dt, "max |norm-1|", float(np.max(np.abs(nrm - 1))), "particles", float(v[-1, :, 0].sum()))
print("<N_j> at t=T:", np.round(v[-1, :, 0], 4))
