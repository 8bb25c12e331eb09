"""Golden vector at the cfg5 shape (BASELINE.json configs[4]: BH L=50 Npart=50 d=5, chi=256, Cutoff 1e-10) made by the oracle (CPU).
Run from the repo root:   python tests/golden/make_golden_cfg5.py

golden_cfg5_quench.npz: the Mott product state |1 1 ... 1> quenched at U = 2.5 with time step 0.05 (bench.py's cfg5 protocol), NQ = 32
steps: bond dimensions after every step (the bulk bonds reach 256 at step ~28), the return amplitude <psi_0|psi_k> after every step.
Pins the truncation decisions of the 256-row charge blocks (cluster kernels) on the spectra of a physical quench."""
import os, sys, time
HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
import numpy as np
import bench
from oracle import bh_mps as ob

NQ = 32


def main():
    c = bench.CFG5
    L, D = c["L"], c["d"] + 1
    st = ob.BHStepper(L, D, c["J"], 5 * c["tstep"], ob.TruncArgs(cutoff=c["cutoff"], maxm=c["maxm"]))
    psi0 = ob.product_state([1] * L, D)
    psi = psi0.copy()
    dims, amp = [psi.bond_dims()], [ob.overlap(psi0, psi)]
    t0 = time.time()
    for k in range(NQ):
        st.step(psi, 2.5, 2.5, True)
        dims.append(psi.bond_dims()); amp.append(ob.overlap(psi0, psi))
        print(k, max(psi.bond_dims()), f"{time.time() - t0:.1f}s", flush=True)
    np.savez_compressed(os.path.join(HERE, "golden_cfg5_quench.npz"), dims=np.array(dims), amp=np.array(amp), NQ=np.array(NQ),
                        cpu_seconds=np.array(time.time() - t0))
    print("cfg5 quench: max dim", int(np.max(dims)), "seconds", time.time() - t0)


if __name__ == "__main__":
    main()
