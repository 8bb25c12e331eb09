"""Measured comparison only (the product path never calls cuSOLVER): cuSOLVER's one-sided Jacobi SVD (gesvdj, through
torch.linalg.svd(driver="gesvdj")) and its QR-iteration SVD (gesvd) on the REAL charge blocks of a saturated cfg2 gate
decomposition, against the engine's own decomposition of the same two-site tensor (merge + gate, block table, pivoted QR + Jacobi
of all blocks concurrently, truncation, factor assembly).
usage: gpu_cusolver_compare.py [K0=170]"""
import sys, os, time, ctypes
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
import bench
import optimalcontrolmps_b200 as oc
from optimalcontrolmps_b200.states import ground_state
from oracle import bh_mps as ob

CFG = bench.CFG
K0 = int(sys.argv[1]) if len(sys.argv) > 1 else 170
L, d = CFG["L"], CFG["d"]
D = d + 1
st = oc.BH_tDMRG(oc.BoseHubbard(L, d), CFG["J"], CFG["tstep"], oc.Args("Cutoff=", CFG["cutoff"], "Maxm=", CFG["maxm"]))
basis, c, u = bench.make_problem_host(0)
psi = st.to_device(ground_state(L, d, CFG["Npart"], CFG["U_i"]))
for k in range(K0):
    st.step(psi, u[k], u[k + 1], True)

# ---- the engine: time the ops of ONE step that belong to the gate on the central bond (sites 9,10 -> first forward half sweep) ----
sched = st.schedule()
lib = st.ctx.lib
site = 9                                                  # gate on sites (9, 10), 1-based; the centre arrives there during the sweep
idx = [i for i, o in enumerate(sched) if o[0] == 1 and o[1] == site][0]
work = st.new_mps().copy_from(psi)
lib.ocmps_debug_run_ops.argtypes = [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_double, ctypes.c_double, ctypes.c_int, ctypes.c_int, ctypes.c_int]
lib.ocmps_debug_run_ops(st.h, work.h, u[K0], u[K0 + 1], 1, 0, idx)           # everything before the gate
host_before = work.download()
ts = []
for rep in range(5):
    w2 = st.new_mps().copy_from(work)
    w2_host_lims = None
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    lib.ocmps_debug_run_ops(st.h, w2.h, u[K0], u[K0 + 1], 1, idx, idx + 1)    # merge+gate, setup, QR+Jacobi, truncate, build
    ts.append(time.perf_counter() - t0)
print(f"engine: gate decomposition on sites ({site},{site + 1}) of step {K0}: {min(ts) * 1e6:.0f} us (best of 5, plain launches, host-synchronous; "
      f"all charge blocks concurrently, including merge+gate, truncation and factor assembly)")

# ---- the same blocks through cuSOLVER ----
po = ob.MPS(host_before.A, [np.asarray(x, dtype=np.int64) for x in host_before.q], site - 1, site + 1)
G = ob.bond_gate(D, CFG["J"], CFG["tstep"]).reshape(D, D, D, D)
u1 = ob.u_phases(D, u[K0], CFG["tstep"])
th = np.tensordot(po.A[site - 1], po.A[site], axes=(2, 0))
th = th * (u1[None, :, None, None] * u1[None, None, :, None])
th = np.einsum("tuab,labr->ltur", G, th, optimize=True)
chil, _, _, chir = th.shape
s = np.arange(D)
X = th.reshape(chil * D, D * chir)
rowq = (po.q[site - 1][:, None] + s[None, :]).ravel()
colq = (po.q[site + 1][None, :] - s[:, None]).ravel()
blocks = []
for q in np.unique(rowq):
    r = np.nonzero(rowq == q)[0]; cc = np.nonzero(colq == q)[0]
    if len(cc):
        blocks.append((int(q), X[np.ix_(r, cc)]))
dev = torch.device("cuda", 0)


def time_svd(a, driver, reps=10):
    t = torch.from_numpy(np.ascontiguousarray(a)).to(dev)
    torch.linalg.svd(t, full_matrices=False, driver=driver)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    best = 1e9
    for _ in range(reps):
        e0.record()
        U, S, Vh = torch.linalg.svd(t, full_matrices=False, driver=driver)
        e1.record()
        torch.cuda.synchronize()
        best = min(best, e0.elapsed_time(e1))
    ref = np.linalg.svd(a, compute_uv=False)
    err = float(np.max(np.abs(S.cpu().numpy() - ref) / ref[0]))
    return best * 1e3, err


tot = {"gesvdj": 0.0, "gesvd": 0.0}
mx = {"gesvdj": 0.0, "gesvd": 0.0}
print("charge  shape      cuSOLVER gesvdj us   gesvd us   (max |sigma - LAPACK| / sigma_0)")
for q, b in blocks:
    row = f"{q:5d}  {str(b.shape):10s}"
    for drv in ("gesvdj", "gesvd"):
        us, err = time_svd(b, drv)
        tot[drv] += us; mx[drv] = max(mx[drv], us)
        row += f"  {us:10.0f} ({err:.0e})"
    print(row)
print(f"cuSOLVER, blocks one after the other: gesvdj {tot['gesvdj']:.0f} us, gesvd {tot['gesvd']:.0f} us; slowest single block: gesvdj {mx['gesvdj']:.0f} us, "
      f"gesvd {mx['gesvd']:.0f} us (a perfectly concurrent batch could not be faster than that)")
