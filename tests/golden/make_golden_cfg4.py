"""Golden vectors at the cfg4 shape (BASELINE.json configs[3]: BH L=30 Npart=30 d=5, chi=150, Cutoff 1e-8) from PHYSICAL states, made by
the oracle (CPU, a few minutes).  Run from the repo root:

    python tests/golden/make_golden_cfg4.py

golden_cfg4_sweep.npz: the first NT4 = 171 time points of bench.py's cfg4 control of seed 1 (GROUP M=10 chopped sines on the
linsigmoid ramp 2.5 -> 50), evaluated as a GRAPE problem of that horizon from the L=30 DMRG ground states (the fixtures in
optimalcontrolmps_b200/data): cost, gradient, divT, the fidelities and the bond dimensions of all slices of psi_t and xi_t.  The bulk bonds
reach chi = 150 along the way, so every truncation decision of the charge blocks that need the cluster kernels (more than 64 rows of R) is
pinned on graded, physical spectra."""
import os, sys, time
HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
import numpy as np
import bench
from oracle import bh_mps as ob, optimal_control as oo

NT4 = 171


def cfg4_control(seed=1):
    import optimalcontrolmps_b200.api as api          # host-only helpers (SeedGenerator / ControlBasisFactory mirrors)
    c = bench.CFG4
    N = int(c["T"] / c["tstep"] + 1)
    u0 = api.SeedGenerator.linsigmoidSeed(c["U_i"], c["U_f"], N, np.random.default_rng(7))
    basis = api.ControlBasisFactory.buildChoppedSineBasis(u0, c["tstep"], c["T"], c["M"])
    cc = np.array(api.SeedGenerator.randomCoeffSeed(-4.0, 4.0, c["M"], np.random.default_rng(4000 + seed)))
    for _ in range(40):
        uu = np.array(basis.convertControl(list(cc)))
        if uu.min() >= 2.0 and uu.max() <= 100.0:
            break
        cc *= 0.8
    return np.array(basis.convertControl(list(cc)))


def main():
    from optimalcontrolmps_b200.states import ground_state
    c = bench.CFG4
    D = c["d"] + 1
    conv = lambda h: ob.MPS(h.A, [np.asarray(x, dtype=np.int64) for x in h.q], 0, 2)
    psi_i = conv(ground_state(c["L"], c["d"], c["Npart"], c["U_i"]))
    psi_f = conv(ground_state(c["L"], c["d"], c["Npart"], c["U_f"]))
    u = cfg4_control(1)[:NT4]
    st = ob.BHStepper(c["L"], D, c["J"], c["tstep"], ob.TruncArgs(cutoff=c["cutoff"], maxm=c["maxm"]))
    oc = oo.OptimalControl(psi_f, psi_i, st, N=NT4, gamma=c["gamma"])
    t0 = time.time()
    out = {"u": u, "NT4": np.array(NT4)}
    out["grad"] = np.array(oc.getAnalyticGradient(list(u), True))
    out["cost"] = np.array(oc.getCost(list(u), False))
    out["fidelities"] = np.array(oc.getFidelityForAllT(list(u), False))
    out["psi_dims"] = np.array([x.bond_dims() for x in oc.psi_t])
    out["xi_dims"] = np.array([x.bond_dims() for x in oc.xi_t])
    out["divT"] = np.array(oc.divT)
    out["cpu_seconds"] = np.array(time.time() - t0)
    np.savez_compressed(os.path.join(HERE, "golden_cfg4_sweep.npz"), **out)
    print("cfg4 sweep: cost", float(out["cost"]), "seconds", float(out["cpu_seconds"]), "max dim", int(out["psi_dims"].max()),
          "slices at 150:", int((out["psi_dims"].max(axis=1) == 150).sum()))


if __name__ == "__main__":
    main()
