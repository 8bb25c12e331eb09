"""Development timing (GPU): cfg2-shaped forward steps."""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import optimalcontrolmps_b200 as oc
from oracle import bh_mps as ob, ground_state as og

L, d, Np = 20, 5, 20
D = d + 1
t0 = time.time()
psi_i = og.ground_state_dmrg(L, D, Np, 1.0, 2.5, maxm_schedule=(10, 20, 50, 100), cutoff=1e-8)
psi_f = og.ground_state_dmrg(L, D, Np, 1.0, 50.0, maxm_schedule=(10, 20, 50, 100), cutoff=1e-8)
print("dmrg", time.time() - t0, psi_i.bond_dims())
st = oc.BH_tDMRG(oc.BoseHubbard(L, d), 1.0, 1e-2, oc.Args("Cutoff=", 1e-8, "Maxm=", 100))
Nt = int(sys.argv[1]) if len(sys.argv) > 1 else 21
u = np.array(oc.SeedGenerator.linsigmoidSeed(2.5, 50, 201, np.random.default_rng(1)))[:Nt] if Nt <= 201 else None
# stretch the ramp so that entanglement is generated quickly: use full ramp sampled coarsely
u = np.array(oc.SeedGenerator.linsigmoidSeed(2.5, 50, Nt, np.random.default_rng(1)))
ocg = oc.OptimalControl(oc.IQMPS(psi_f.A, psi_f.q), oc.IQMPS(psi_i.A, psi_i.q), st, Nt, 1e-6)
ocg.setThreadCount(2)
for rep in range(2):
    t0 = time.time(); c = ocg.getCost(list(u)); t1 = time.time()
    print("cost", c, "time", t1 - t0, "per step", (t1 - t0) / (Nt - 1))
t0 = time.time(); g = ocg.getAnalyticGradient(list(u)); t1 = time.time()
print("grad time", t1 - t0, "per step-pair", (t1 - t0) / (Nt - 1))
print(ocg.psi_t.bond_dims()[-1].tolist())
print(ocg.xi_t.bond_dims()[0].tolist())
# oracle timing for a few steps from the last slice
po = ob.MPS(*[(h.A, [x.astype(np.int64) for x in h.q]) for h in [ocg.psi_t.get(Nt - 1).download()]][0])
so = ob.BHStepper(L, D, 1.0, 1e-2, ob.TruncArgs(cutoff=1e-8, maxm=100))
t0 = time.time()
for k in range(2): so.step(po, 30.0, 30.0, True)
print("oracle per step", (time.time() - t0) / 2, po.bond_dims())
