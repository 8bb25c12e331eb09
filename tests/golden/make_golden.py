"""Generates the golden input/output vectors under tests/golden/ from the oracle.

The reference cannot be built here (ITensor / IPOPT / GoogleTest are absent), so
these vectors come from the NumPy restatement in oracle/, which itself is pinned to the reference's
own goldens (tests/test_oracle_goldens.py).  Run from the repo root:

    python tests/golden/make_golden.py
"""
import os, sys
HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
import numpy as np
from oracle import bh_mps as ob, optimal_control as oo, ground_state as og


def pack_state(prefix, psi, out):
    for j in range(psi.L):
        out[f"{prefix}_A{j}"] = psi.A[j]
    for b in range(psi.L + 1):
        out[f"{prefix}_q{b}"] = psi.q[b].astype(np.int32)


def load_fixture(name):
    z = np.load(os.path.join(os.path.dirname(os.path.dirname(HERE)), "optimalcontrolmps_b200", "data", name))
    L = sum(1 for k in z.files if k.startswith("A"))
    return ob.MPS([z[f"A{j}"] for j in range(L)], [z[f"q{b}"].astype(np.int64) for b in range(L + 1)])


def case(name, L, d, Npart, J, cs, ce, T, ts, cutoff, maxm, M, seed, states=None, gamma=1e-6, hessian=True):
    D = d + 1
    N = int(T / ts + 1)
    if states is None:
        psi_i = og.ground_state_ed(L, D, Npart, J, cs)
        psi_f = og.ground_state_ed(L, D, Npart, J, ce)
    else:
        psi_i, psi_f = states
    args = ob.TruncArgs(cutoff=cutoff, maxm=maxm)
    st = ob.BHStepper(L, D, J, ts, args)
    rng = np.random.default_rng(seed)
    out = {"params": np.array([L, d, Npart, J, cs, ce, T, ts, cutoff, -1 if maxm is None else maxm, M, gamma, N], dtype=float)}
    pack_state("init", psi_i, out)
    pack_state("target", psi_f, out)
    # one forward and one backward step
    p = psi_i.copy(); st.step(p, 3.0, 4.5, True); pack_state("step_fwd", p, out)
    p = psi_f.copy(); st.step(p, 7.0, 6.0, False); pack_state("step_bwd", p, out)
    # GRAPE
    u = rng.uniform(2.0, 10.0, N)
    oc = oo.OptimalControl(psi_f, psi_i, st, N=N, gamma=gamma)
    out["u"] = u
    out["cost"] = np.array(oc.getCost(list(u)))
    out["fidelities"] = np.array(oc.getFidelityForAllT(list(u), False))
    out["grad"] = np.array(oc.getAnalyticGradient(list(u)))
    out["psi_dims"] = np.array([x.bond_dims() for x in oc.psi_t])
    out["xi_dims"] = np.array([x.bond_dims() for x in oc.xi_t])
    out["divT"] = np.array(oc.divT)
    if hessian:
        out["hessian"] = np.array(oc.getHessian(list(u)))
        kp = ob.apply_K(oc.psi_t[N // 2], args)
        out["applyK_dims"] = np.array(kp.bond_dims())
        out["applyK_norm"] = np.array(kp.norm())
        out["applyK_ovl"] = np.array(ob.overlap(oc.psi_t[N // 2], kp))
    # GROUP
    u0 = oo.linspace(cs, ce, N)
    basis = oo.build_chopped_sine_basis(u0, ts, T, M)
    og_ = oo.OptimalControl(psi_f, psi_i, st, basis=basis, gamma=gamma)
    c = rng.uniform(-1.0, 1.0, M)
    out["c"] = c
    out["group_cost"] = np.array(og_.getCost(list(c)))
    out["group_grad"] = np.array(og_.getAnalyticGradient(list(c)))
    if hessian:
        out["group_hessian"] = np.array(og_.getHessian(list(c)))
    np.savez_compressed(os.path.join(HERE, name), **out)
    print(name, "cost", float(out["cost"]), "dims", out["psi_dims"][-1].tolist())


if __name__ == "__main__":
    case("golden_L5.npz", 5, 5, 5, 1.0, 2.0, 12.0, 0.1, 1e-2, 1e-8, None, 4, seed=11)
    case("golden_L6_maxm.npz", 6, 4, 6, 1.0, 2.5, 20.0, 0.12, 1e-2, 1e-8, 12, 3, seed=12)
    case("golden_L3.npz", 3, 3, 3, 2.0, 2.0, 12.0, 0.2, 1e-2, 1e-7, None, 3, seed=13)
    s = (load_fixture("bh_L8_d4_N8_U2.5.npz"), load_fixture("bh_L8_d4_N8_U50.npz"))
    case("golden_L8_maxm.npz", 8, 4, 8, 1.0, 2.5, 50.0, 0.2, 1e-2, 1e-8, 24, 5, seed=14, states=s, hessian=True)
