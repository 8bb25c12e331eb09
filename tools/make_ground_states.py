"""Generates the Bose-Hubbard ground-state fixtures used as psi_init / psi_target by bench.py and the
full-size GPU tests (reference: include/InitializeState.hpp:69-117 with maxBondDim / threshold, as
main/OptimizeRamp.cpp:85-86 calls it).  Uses the oracle's two-site DMRG; run once, fixtures are committed.

    python tools/make_ground_states.py
"""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from oracle import ground_state as og

OUT = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "optimalcontrolmps_b200", "data")


def save(name, psi, meta):
    arrs = {f"A{j}": psi.A[j] for j in range(psi.L)}
    arrs.update({f"q{b}": psi.q[b].astype(np.int32) for b in range(psi.L + 1)})
    arrs["meta"] = np.array(meta, dtype=np.float64)
    np.savez_compressed(os.path.join(OUT, name), **arrs)


def main():
    configs = [  # (L, d, Npart, maxm, cutoff)
        (20, 5, 20, 100, 1e-8),     # cfg2 / cfg3
        (8, 4, 8, 40, 1e-9),        # mid-size test problem
        (5, 4, 5, 0, 0.0),          # cfg1 (README input): exact diagonalisation
        (30, 5, 30, 150, 1e-8),     # cfg4 (batched seeds)
    ]
    only = [int(x) for x in sys.argv[1:]]          # optional: chain lengths to (re)generate
    for (L, d, Np, maxm, cutoff) in configs:
        if only and L not in only:
            continue
        for U in (2.5, 50.0):
            t0 = time.time()
            if maxm == 0:
                psi = og.ground_state_ed(L, d + 1, Np, 1.0, U)
            else:
                psi = og.ground_state_dmrg(L, d + 1, Np, 1.0, U, maxm_schedule=(10, 20, 50, maxm), cutoff=cutoff, nsweeps=10)
            e = og.mps_energy(psi, 1.0, U)
            print(f"L={L} d={d} U={U}: E={e:.10f} dims={psi.bond_dims()} ({time.time() - t0:.1f}s)", flush=True)
            save(f"bh_L{L}_d{d}_N{Np}_U{U:g}.npz", psi, [L, d, Np, 1.0, U, maxm, cutoff, e])


if __name__ == "__main__":
    main()
