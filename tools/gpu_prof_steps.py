"""Profiling workload: evolve a cfg2-shaped state for a few steps (run under ncu)."""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import optimalcontrolmps_b200 as oc
from oracle import ground_state as og

L, d, Np = 20, 5, 20
D = d + 1
psi_i = og.ground_state_dmrg(L, D, Np, 1.0, 2.5, maxm_schedule=(10, 20, 50, 100), cutoff=1e-8)
st = oc.BH_tDMRG(oc.BoseHubbard(L, d), 1.0, 1e-2, oc.Args("Cutoff=", 1e-8, "Maxm=", 100))
nwarm = int(sys.argv[1]) if len(sys.argv) > 1 else 30
nprof = int(sys.argv[2]) if len(sys.argv) > 2 else 2
pd = st.to_device(oc.IQMPS(psi_i.A, psi_i.q))
u = np.linspace(2.5, 50, nwarm + nprof + 1)
lib = st.ctx.lib
c0 = lib.ocmps_launch_count()
for k in range(nwarm):
    st.step(pd, u[k], u[k + 1], True)
c1 = lib.ocmps_launch_count()
dbg0=(ctypes.c_ulonglong*8)() if False else None
import ctypes as _c
_d=(_c.c_ulonglong*8)(); lib.ocmps_debug_jacobi(_d,1)
t0 = time.time()
for k in range(nwarm, nwarm + nprof):
    st.step(pd, u[k], u[k + 1], True)
t1 = time.time()
c2 = lib.ocmps_launch_count()
import ctypes
dbg=(ctypes.c_ulonglong*8)()
lib.ocmps_debug_jacobi(dbg,0)
print("jacobi dbg: sweeps",dbg[0],"blocks",dbg[1],"max sweeps",dbg[2],"big-block sweeps",dbg[3],"big blocks",dbg[4], "per big block: QR kclk",dbg[5]/max(dbg[4],1)/1e3,"Jacobi kclk",dbg[6]/max(dbg[4],1)/1e3,"total kclk",dbg[7]/max(dbg[4],1)/1e3)
print("warm launches", c1 - c0, "prof launches", c2 - c1, "per step ms", (t1 - t0) / nprof * 1e3, pd.bond_dims())
