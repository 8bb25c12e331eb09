"""Time Trotter steps at the shapes of BASELINE.json configs[3] / configs[4] (random number-conserving MPS, every bond saturated)."""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import numpy as np
import ctypes
import optimalcontrolmps_b200 as oc
from conftest import random_symmetric_mps, to_host

for (L, Np, chi, cutoff) in [(30, 30, 150, 1e-8), (50, 50, 256, 1e-10)]:
    psi = random_symmetric_mps(L, 6, Np, chi, seed=L)
    st = oc.BH_tDMRG(oc.BoseHubbard(L, 5), 1.0, 1e-2, oc.Args("Cutoff=", cutoff, "Maxm=", chi))
    dev = st.to_device(to_host(psi))
    lib = st.ctx.lib
    u = np.linspace(3.0, 4.0, 12)
    for k in range(3):
        st.step(dev, u[k], u[k + 1], True)
    dev.norm()
    d = (ctypes.c_ulonglong * 8)(); lib.ocmps_debug_jacobi(d, 1)
    t0 = time.time()
    n = 4
    for k in range(3, 3 + n):
        st.step(dev, u[k], u[k + 1], True)
    dev.norm()
    dt = (time.time() - t0) / n
    lib.ocmps_debug_jacobi(d, 0)
    print("L", L, "chi", chi, "cutoff", cutoff, "ms/step", dt * 1e3, "max sweeps", d[2], "blocks/step", d[1] / n, "dims", max(dev.bond_dims()), flush=True)
