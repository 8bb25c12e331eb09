import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, ctypes as C
import optimalcontrolmps_b200 as oc
from optimalcontrolmps_b200 import _lib
from optimalcontrolmps_b200.api import _pd, _pi
from optimalcontrolmps_b200.states import ground_state
L, d = 20, 5
st = oc.BH_tDMRG(oc.BoseHubbard(L, d), 1.0, 1e-2, oc.Args("Cutoff=", 1e-8, "Maxm=", 100))
psi_i, psi_f = ground_state(L, d, 20, 2.5), ground_state(L, d, 20, 50.0)
Nt = int(sys.argv[1]) if len(sys.argv) > 1 else 71
u = np.linspace(2.5, 30.0, Nt)
o = oc.OptimalControl(psi_f, psi_i, st, Nt, 1e-6)
o.setThreadCount(2)
lib = o.lib
for rep in range(2):
    t0 = time.time(); o._calcPsiXiDivT(list(u)); t1 = time.time()
    if o.xiHlist is None: o.xiHlist = st.new_store(Nt)
    _lib.check(lib.ocmps_store_apply_K(st.h, o.xi_t.h, Nt, o.xiHlist.h)); t2 = time.time()
    rows = np.arange(1, Nt - 1, dtype=np.int32); ovl = np.zeros(2 * Nt * Nt); norms = np.zeros(Nt)
    _lib.check(lib.ocmps_hessian_rows(st.h, o.psi_t.h, o.xiHlist.h, _pd(u), Nt, _pi(rows), rows.size, int(sys.argv[2]) if len(sys.argv) > 2 else 48, _pd(ovl), _pd(norms))); t3 = time.time()
    print("sweeps+divT %.2f  store_apply_K %.2f  rows %.2f  (row-steps %d)" % (t1 - t0, t2 - t1, t3 - t2, (Nt - 2) * (Nt - 3) // 2))
x = o.psi_t.get(Nt // 2)
t0 = time.time()
for k in range(5): y = st.exactApplyMPO(x)
print("single apply_K %.3f s" % ((time.time() - t0) / 5), y.bond_dims())
