"""Helper of test_gpu_switches.py: evaluates cost / gradient / Hessian of one golden problem on the GPU and prints them as
JSON.  Run in a subprocess because the engine reads its expert switches (OCMPS_*) once per process."""
import json
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from conftest import golden_state, load_golden, to_host  # noqa: E402


def main():
    import optimalcontrolmps_b200 as oc
    z = load_golden(sys.argv[1])
    L, d, Np, J, cs, ce, T, ts, cutoff, maxm, M, gamma, N = z["params"]
    L, d, N = int(L), int(d), int(N)
    maxm = None if maxm < 0 else int(maxm)
    init, target = golden_state(z, "init"), golden_state(z, "target")
    a = oc.Args("Cutoff=", cutoff) if maxm is None else oc.Args("Cutoff=", cutoff, "Maxm=", maxm)
    cap = None if maxm is None else max([maxm] + [max(s.bond_dims()) for s in (init, target)])
    st = oc.BH_tDMRG(oc.BoseHubbard(L, d), J, ts, a, chi_cap=cap)
    u = list(z["u"])
    g = oc.OptimalControl(to_host(target), to_host(init), st, N, gamma)
    g.setThreadCount(2)
    grad = np.array(g.getAnalyticGradient(u, True))
    cost = g.getCost(u, False)
    H = np.array(g.getHessian(u, False))
    print("PROBE " + json.dumps({"cost": cost, "grad": grad.tolist(), "hess": H.tolist(),
                                 "dims": np.asarray(g.psi_t.bond_dims()).tolist()}))


if __name__ == "__main__":
    main()
