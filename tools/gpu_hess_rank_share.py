"""Times what ONE rank of an 8-GPU sharded Hessian does (prerequisites + its rows), on one GPU."""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import optimalcontrolmps_b200 as oc
from optimalcontrolmps_b200 import _lib, distributed
from optimalcontrolmps_b200.api import _pd, _pi
from optimalcontrolmps_b200.states import ground_state
L, d = 20, 5
st = oc.BH_tDMRG(oc.BoseHubbard(L, d), 1.0, 1e-2, oc.Args("Cutoff=", 1e-8, "Maxm=", 100))
psi_i, psi_f = ground_state(L, d, 20, 2.5), ground_state(L, d, 20, 50.0)
Nt = 201
world = int(sys.argv[1]) if len(sys.argv) > 1 else 8
chains = int(sys.argv[2]) if len(sys.argv) > 2 else 16
u = np.linspace(2.5, 30.0, Nt)
o = oc.OptimalControl(psi_f, psi_i, st, Nt, 1e-6)
o.setThreadCount(2)
lib = o.lib
t0 = time.time(); o._calcPsiXiDivT(list(u)); t1 = time.time()
o.xiHlist = st.new_store(Nt)
_lib.check(lib.ocmps_store_apply_K(st.h, o.xi_t.h, Nt, o.xiHlist.h)); t2 = time.time()
for rep in range(2):
    rows = np.array(distributed.partition_rows(Nt, world, 0), dtype=np.int32); ovl = np.zeros(2 * Nt * Nt); norms = np.zeros(Nt)
    t2 = time.time()
    _lib.check(lib.ocmps_hessian_rows(st.h, o.psi_t.h, o.xiHlist.h, _pd(u), Nt, _pi(rows), rows.size, chains, _pd(ovl), _pd(norms))); t3 = time.time()
    print("world %d chains %d: sweeps+divT %.2f  rows %.2f  (%d rows, %d row-steps, longest %d)" % (world, chains, t1 - t0, t3 - t2, rows.size, sum(Nt - 2 - r for r in rows), Nt - 2 - rows.min()))
