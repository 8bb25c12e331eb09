"""Round-2 golden vectors from the oracle (CPU, several minutes).  Run from the repo root:

    python tests/golden/make_golden_round2.py [cfg1] [cfg3] [cfg5]

* golden_cfg1.npz -- BASELINE.json configs[0], the README input exactly: BH L=5 Npart=5 d=4, tstep=0.01, T=2.0 (Nt=201),
  Maxm=80, Cutoff=1e-8, GRAPE control = linsigmoid ramp 2.5 -> 50 + uniform(-0.5, 0.5) noise (seeded): cost, gradient,
  fidelities of all 201 slices, bond dimensions (README.md:30-45, main/OptimizeRamp.cpp:36-38).
* golden_cfg3_hessian.npz -- the cfg3 SHAPE (L=20, d=5, chi=100, GROUP M=20) on a reduced horizon: the start state is the
  psi of bench.py's cfg2 control after K0=150 Trotter steps (bond dimensions saturated at 100), evolved by the ORACLE; the
  horizon is the next Nt=13 time points of the same control.  GRAPE Hessian (13x13), GROUP Hessian (20x20), cost, gradient,
  bond dimensions of all slices.  The GPU test rebuilds the start state with the engine itself (the two agree to ~1e-11).
* golden_cfg5_dims.npz -- cfg5 truncation semantics: L=20 ground state (U=2.5), Cutoff=1e-10, Maxm=256, 24 forward steps of a
  steep ramp: bond dimensions of every slice and the final norm/overlap data.
"""
import os, sys, time
HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
import numpy as np
import bench
from oracle import bh_mps as ob, optimal_control as oo, ground_state as og

K0, NT3, M3 = 150, 13, 20


def cfg1():
    L, d, Np, J, ts, T, maxm, cutoff, gamma = 5, 4, 5, 1.0, 0.01, 2.0, 80, 1e-8, 1e-6
    D = d + 1
    N = int(T / ts + 1)
    psi_i = og.ground_state_ed(L, D, Np, J, 2.5)
    psi_f = og.ground_state_ed(L, D, Np, J, 50.0)
    st = ob.BHStepper(L, D, J, ts, ob.TruncArgs(cutoff=cutoff, maxm=maxm))
    import optimalcontrolmps_b200.api as api          # host-only helper: the seed generator mirrors include/SeedGenerator.hpp
    u0 = np.array(api.SeedGenerator.linsigmoidSeed(2.5, 50.0, N, np.random.default_rng(7)))
    u = u0 + np.random.default_rng(2024).uniform(-0.5, 0.5, N)
    u = np.clip(u, 2.0, 100.0)
    oc = oo.OptimalControl(psi_f, psi_i, st, N=N, gamma=gamma)
    out = {"params": np.array([L, d, Np, J, 2.5, 50.0, T, ts, cutoff, maxm, 0, gamma, N], dtype=float), "u": u}
    for j in range(L):
        out[f"init_A{j}"] = psi_i.A[j]; out[f"target_A{j}"] = psi_f.A[j]
    for b in range(L + 1):
        out[f"init_q{b}"] = psi_i.q[b].astype(np.int32); out[f"target_q{b}"] = psi_f.q[b].astype(np.int32)
    t0 = time.time()
    out["grad"] = np.array(oc.getAnalyticGradient(list(u), True))
    out["cost"] = np.array(oc.getCost(list(u), False))
    out["fidelities"] = np.array(oc.getFidelityForAllT(list(u), False))
    out["psi_dims"] = np.array([x.bond_dims() for x in oc.psi_t])
    out["xi_dims"] = np.array([x.bond_dims() for x in oc.xi_t])
    out["divT"] = np.array(oc.divT)
    out["cpu_seconds"] = np.array(time.time() - t0)
    np.savez_compressed(os.path.join(HERE, "golden_cfg1.npz"), **out)
    print("cfg1 cost", float(out["cost"]), "seconds", float(out["cpu_seconds"]), "max dim", int(out["psi_dims"].max()))


def cfg3():
    CFG = bench.CFG
    basis_p, c2, u = bench.make_problem_host(0)
    psi_i, psi_f = bench.oracle_states()
    D = CFG["d"] + 1
    st = ob.BHStepper(CFG["L"], D, CFG["J"], CFG["tstep"], ob.TruncArgs(cutoff=CFG["cutoff"], maxm=CFG["maxm"]))
    p = psi_i.copy()
    t0 = time.time()
    for k in range(K0):
        st.step(p, u[k], u[k + 1], True)
    print("evolved", K0, "steps", time.time() - t0, p.bond_dims(), flush=True)
    uw = np.array(u[K0:K0 + NT3])                       # the horizon: the next NT3 points of the control
    T = (NT3 - 1) * CFG["tstep"]
    out = {"K0": np.array(K0), "u_full": np.array(u), "u": uw, "start_dims": np.array(p.bond_dims())}
    og_ = oo.OptimalControl(psi_f, p, st, N=NT3, gamma=CFG["gamma"])
    t0 = time.time()
    out["hessian"] = np.array(og_.getHessian(list(uw), True))
    print("GRAPE hessian", time.time() - t0, flush=True)
    out["cost"] = np.array(og_.getCost(list(uw), False))
    out["grad"] = np.array(og_.getAnalyticGradient(list(uw), False))
    out["psi_dims"] = np.array([x.bond_dims() for x in og_.psi_t])
    out["xi_dims"] = np.array([x.bond_dims() for x in og_.xi_t])
    out["xiH_dims"] = np.array([x.bond_dims() for x in og_.xiHlist])
    basis = oo.build_chopped_sine_basis(list(uw), CFG["tstep"], T, M3)
    c = np.random.default_rng(33).uniform(-2.0, 2.0, M3)
    ogg = oo.OptimalControl(psi_f, p, st, basis=basis, gamma=CFG["gamma"])
    t0 = time.time()
    out["c"] = c
    out["group_hessian"] = np.array(ogg.getHessian(list(c), True))
    out["group_cost"] = np.array(ogg.getCost(list(c), False))
    out["group_grad"] = np.array(ogg.getAnalyticGradient(list(c), False))
    print("GROUP hessian", time.time() - t0, flush=True)
    np.savez_compressed(os.path.join(HERE, "golden_cfg3_hessian.npz"), **out)
    print("cfg3 cost", float(out["cost"]), "max |H|", float(np.abs(out["hessian"]).max()))


def cfg5():
    CFG = bench.CFG
    psi_i, _ = bench.oracle_states()
    D = CFG["d"] + 1
    st = ob.BHStepper(CFG["L"], D, CFG["J"], CFG["tstep"], ob.TruncArgs(cutoff=1e-10, maxm=256))
    nsteps = 24
    u = np.linspace(2.5, 40.0, nsteps + 1)
    p = psi_i.copy()
    dims = [p.bond_dims()]
    t0 = time.time()
    for k in range(nsteps):
        st.step(p, u[k], u[k + 1], True)
        dims.append(p.bond_dims())
    out = {"u": u, "dims": np.array(dims), "norm": np.array(p.norm()), "ovl0": np.array(ob.overlap(psi_i, p))}
    np.savez_compressed(os.path.join(HERE, "golden_cfg5_dims.npz"), **out)
    print("cfg5", time.time() - t0, "s, final dims", dims[-1])


if __name__ == "__main__":
    which = sys.argv[1:] or ["cfg1", "cfg5", "cfg3"]
    for w in which:
        {"cfg1": cfg1, "cfg3": cfg3, "cfg5": cfg5}[w]()
