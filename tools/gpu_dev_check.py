"""Development check (GPU): step / cost / gradient / Hessian of libocmps against the oracle."""
import sys, time, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import optimalcontrolmps_b200 as oc
from oracle import bh_mps as ob, optimal_control as oo, ground_state as og


def to_host(m):
    return oc.IQMPS(m.A, m.q, m.llim, m.rlim)


def to_oracle(h):
    return ob.MPS(h.A, [x.astype(np.int64) for x in h.q], h.llim, h.rlim)


def run(L, d, Npart, J, cs, ce, T, ts, cutoff, maxm, hess=True, seed=1):
    D = d + 1
    N = int(T / ts + 1)
    psi_i = og.ground_state_ed(L, D, Npart, J, cs)
    psi_f = og.ground_state_ed(L, D, Npart, J, ce)
    args_o = ob.TruncArgs(cutoff=cutoff, maxm=maxm)
    st_o = ob.BHStepper(L, D, J, ts, args_o)
    a = oc.Args("Cutoff=", cutoff) if maxm is None else oc.Args("Cutoff=", cutoff, "Maxm=", maxm)
    cap = None if maxm is None else max([maxm] + psi_i.bond_dims() + psi_f.bond_dims())
    st = oc.BH_tDMRG(oc.BoseHubbard(L, d), J, ts, a, chi_cap=cap)
    print("gate diff", np.abs(st.gate(True) - st_o.G_fwd).max(), np.abs(st.gate(False) - st_o.G_bwd).max())
    # single steps
    rng = np.random.default_rng(seed)
    po = psi_i.copy()
    pd = st.to_device(to_host(psi_i))
    for k in range(3):
        u0, u1 = rng.uniform(2, 10, 2)
        st_o.step(po, u0, u1, True)
        st.step(pd, u0, u1, True)
        h = to_oracle(pd.download())
        print(" step", k, "dims", h.bond_dims(), po.bond_dims(), "|<o|g>|-1", abs(ob.overlap(po, h)) - 1,
              "dense diff", np.abs(po.to_dense() - h.to_dense()).max() if L <= 6 else None, "charge viol", h.check_charges())
    u = list(rng.uniform(2, 10, N))
    oco = oo.OptimalControl(psi_f, psi_i, st_o, N=N, gamma=1e-3)
    ocg = oc.OptimalControl(to_host(psi_f), to_host(psi_i), st, N, 1e-3)
    t0 = time.time(); co = oco.getCost(u); t1 = time.time(); cg = ocg.getCost(u); t2 = time.time()
    print("cost", co, cg, "rel", abs(co - cg) / abs(co), "t oracle", t1 - t0, "t gpu", t2 - t1)
    fo = np.array(oco.getFidelityForAllT(u, False)); fg = np.array(ocg.getFidelityForAllT(u, False))
    print("fid max diff", np.abs(fo - fg).max())
    dims_g = ocg.psi_t.bond_dims(); dims_o = np.array([p.bond_dims() for p in oco.psi_t])
    print("bond dims equal", np.array_equal(dims_g, dims_o), dims_g[-1].tolist())
    t0 = time.time(); go = np.array(oco.getAnalyticGradient(u)); t1 = time.time(); gg = np.array(ocg.getAnalyticGradient(u)); t2 = time.time()
    print("grad rel", np.abs(go - gg).max() / np.abs(go).max(), "t oracle", t1 - t0, "t gpu", t2 - t1)
    ocg.setThreadCount(2)
    gg2 = np.array(ocg.getAnalyticGradient(u))
    print("grad threads=2 vs 1", np.abs(gg2 - gg).max())
    ocg.setBFGS(True); gb = np.array(ocg.getAnalyticGradient(u)); ocg.setBFGS(False)
    oco.setBFGS(True); gbo = np.array(oco.getAnalyticGradient(u)); oco.setBFGS(False)
    print("grad BFGS rel", np.abs(gb - gbo).max() / np.abs(gbo).max())
    if hess:
        # apply_K check
        xo = ob.apply_K(oco.psi_t[N // 2], args_o)
        xg = to_oracle(st.exactApplyMPO(ocg.psi_t.get(N // 2)).download())
        print("applyK dims", xo.bond_dims(), xg.bond_dims(), "norms", xo.norm(), xg.norm(),
              "ovl", abs(ob.overlap(xo, xg)) / (xo.norm() * xg.norm()) - 1)
        t0 = time.time(); Ho = np.array(oco.getHessian(u)); t1 = time.time(); Hg = np.array(ocg.getHessian(u)); t2 = time.time()
        print("hess rel", np.abs(Ho - Hg).max() / np.abs(Ho).max(), "t oracle", t1 - t0, "t gpu", t2 - t1)
    print("launches", oc.Context.default().lib.ocmps_launch_count())


if __name__ == "__main__":
    run(5, 5, 5, 1.0, 2.0, 12.0, 0.1, 1e-2, 1e-8, None)
    run(6, 4, 6, 1.0, 2.5, 20.0, 0.1, 1e-2, 1e-8, 12)
    run(5, 5, 5, 1.0, 2.0, 50.0, 0.2, 1e-2, 1e-8, 10, hess=True, seed=3)
