"""Golden vector at BASELINE.json's full cfg2 size (L=20, Npart=20, d=5, T=2, tstep=0.01, GROUP M=10, Maxm=100,
Cutoff=1e-8): one complete cost+gradient evaluation by the oracle on bench.py's synthetic control (seed 0).
Takes ~2 minutes of CPU.  Run from the repo root:  python tests/golden/make_golden_cfg2.py"""
import os, sys
HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
import numpy as np
import bench
from oracle import bh_mps as ob, optimal_control as oo

basis_p, c, u = bench.make_problem_host(0)
psi_i, psi_f = bench.oracle_states()
CFG = bench.CFG
st = ob.BHStepper(CFG["L"], CFG["d"] + 1, CFG["J"], CFG["tstep"], ob.TruncArgs(cutoff=CFG["cutoff"], maxm=CFG["maxm"]))
basis = oo.ControlBasis(list(basis_p._u0), list(basis_p._S), basis_p._f.tolist())
ocp = oo.OptimalControl(psi_f, psi_i, st, basis=basis, gamma=CFG["gamma"])
grad = np.array(ocp.getAnalyticGradient(list(c), True))
cost = ocp.getCost(list(c), False)
fid = np.array(ocp.getFidelityForAllT(list(c), False))
np.savez_compressed(os.path.join(HERE, "golden_cfg2_eval.npz"), c=c, u=u, cost=np.array(cost), grad=grad, fidelities=fid,
                    divT=np.array(ocp.divT), psi_dims=np.array([p.bond_dims() for p in ocp.psi_t]),
                    xi_dims=np.array([p.bond_dims() for p in ocp.xi_t]))
print("cost", cost, "grad", grad)
