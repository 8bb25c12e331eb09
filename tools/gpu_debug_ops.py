import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, ctypes as C
import optimalcontrolmps_b200 as oc
from optimalcontrolmps_b200 import _lib
from oracle import bh_mps as ob, ground_state as og

L, d, Np = 5, 5, 5
D = d + 1
psi_i = og.ground_state_ed(L, D, Np, 1.0, 2.0)
st = oc.BH_tDMRG(oc.BoseHubbard(L, d), 1.0, 1e-2, oc.Args("Cutoff=", 1e-8))
lib = st.ctx.lib
lib.ocmps_debug_run_ops.argtypes = [C.c_void_p, C.c_void_p, C.c_double, C.c_double, C.c_int, C.c_int, C.c_int]
pd = st.to_device(oc.IQMPS(psi_i.A, psi_i.q, 0, 2))
ops = st.schedule()
print(ops)
prev = psi_i.to_dense()
for k, op in enumerate(ops):
    rc = lib.ocmps_debug_run_ops(st.h, pd.h, 3.0, 4.0, 1, k, k + 1)
    if rc: print("op", k, op, "rc", rc, lib.ocmps_last_error())
    h = pd.download()
    m = ob.MPS(h.A, [x.astype(np.int64) for x in h.q])
    dense = m.to_dense()
    print(k, op, "dims", m.bond_dims(), "norm", np.linalg.norm(dense), "nan", np.isnan(dense).any(),
          "change", np.linalg.norm(dense - prev), "viol", m.check_charges())
    prev = dense
    if rc: break
