"""Device ground-state generator (ocmps_ground_state, imaginary-time evolution with the Trotter-step kernels) against the DMRG
fixtures of optimalcontrolmps_b200/data (oracle two-site DMRG): energy, fidelity, bond dimensions, wall time.
usage: gpu_ground_state.py [tau_final]"""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import optimalcontrolmps_b200 as oc
from optimalcontrolmps_b200.states import ground_state, DATA_DIR

tau = float(sys.argv[1]) if len(sys.argv) > 1 else 0.0
for (L, d, Np, maxm, cutoff) in ((8, 4, 8, 40, 1e-9), (20, 5, 20, 100, 1e-8), (5, 4, 5, 80, 1e-9)):
    for U in (2.5, 50.0):
        ref = ground_state(L, d, Np, U)
        meta = np.load(os.path.join(DATA_DIR, f"bh_L{L}_d{d}_N{Np}_U{U:g}.npz"))["meta"]
        t0 = time.perf_counter()
        psi, e, n = oc.InitializeState(oc.BoseHubbard(L, d), Np, 1.0, U, maxm, cutoff, tau_final=tau, return_info=True)
        dt = time.perf_counter() - t0
        dev_ref = oc.DeviceMPS(psi.ctx, L, d + 1, psi.chi_cap).upload(ref)
        ov = oc.overlapC(dev_ref, psi)
        print(f"L={L} d={d} U={U}: E={e:.10f} (DMRG fixture {meta[7]:.10f}, diff {e - meta[7]:.2e}) 1-|<dmrg|psi>|^2={1 - abs(ov) ** 2:.2e} "
              f"steps {n} time {dt:.2f}s dims max {max(psi.bond_dims())} (fixture {max(ref.bond_dims())}) norm {psi.norm():.12f}", flush=True)
