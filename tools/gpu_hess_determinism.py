import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import optimalcontrolmps_b200 as oc
from optimalcontrolmps_b200.states import ground_state
L, d = 20, 5
st = oc.BH_tDMRG(oc.BoseHubbard(L, d), 1.0, 1e-2, oc.Args("Cutoff=", 1e-8, "Maxm=", 100))
psi_i, psi_f = ground_state(L, d, 20, 2.5), ground_state(L, d, 20, 50.0)
Nt = int(sys.argv[1]) if len(sys.argv) > 1 else 31
u = list(np.linspace(2.5, 30.0, Nt))
res = {}
for (threads, chains) in [(1, 16), (1, 16), (1, 24), (1, 32), (1, 48), (1, 64)]:
    os.environ["OCMPS_HESSIAN_THREADS"] = str(threads)
    o = oc.OptimalControl(psi_f, psi_i, st, Nt, 1e-6)
    o.hessian_chains = chains
    t0 = time.time()
    H = np.array(o.getHessian(u, True))
    dt = time.time() - t0
    key = (threads, chains)
    print(key, "time %.2f" % dt, "checksum", np.abs(H).sum(), "diff vs first", 0.0 if not res else np.abs(H - res["first"]).max())
    if not res:
        res["first"] = H
