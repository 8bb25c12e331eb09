"""cfg5 shape (L=50, chi=256, Cutoff 1e-10): the Mott state quenched until the bonds saturate, then timed forward steps with the
block-SVD cycle counters (QR / Jacobi clocks per block with >= 64 vectors).  OCMPS_STEP_TRACE=1 adds the per-phase timeline."""
import sys, os, time, ctypes
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import bench
import optimalcontrolmps_b200 as oc

c = bench.CFG5
L, d = c["L"], c["d"]
D = d + 1
chi = int(sys.argv[1]) if len(sys.argv) > 1 else c["maxm"]
args = oc.Args("Cutoff=", c["cutoff"], "Maxm=", chi)
quench = oc.BH_tDMRG(oc.BoseHubbard(L, d), c["J"], 5 * c["tstep"], args)
st = oc.BH_tDMRG(oc.BoseHubbard(L, d), c["J"], c["tstep"], args)
A = []
for j in range(L):
    a = np.zeros((1, D, 1), dtype=np.complex128); a[0, 1, 0] = 1.0; A.append(a)
q = [np.array([b], dtype=np.int32) for b in range(L + 1)]
psi = quench.to_device(oc.IQMPS(A, q, 0, 2))
n = 0
while max(psi.bond_dims()) < chi and n < 400:
    quench.step(psi, 2.5, 2.5, True); n += 1
print("quench steps", n, "dims", psi.bond_dims())
lib = st.ctx.lib
for k in range(2):
    st.step(psi, 2.5, 2.6, True)
dbg = (ctypes.c_ulonglong * 8)()
lib.ocmps_debug_jacobi(dbg, 1)
K = int(os.environ.get("OCMPS_TOOL_STEPS", "3"))
rt = ctypes.CDLL("/usr/local/cuda/lib64/libcudart.so")
rt.cudaProfilerStart()
t0 = time.perf_counter()
for k in range(K):
    st.step(psi, 2.6, 2.7, True)
dt = (time.perf_counter() - t0) / K
rt.cudaProfilerStop()
lib.ocmps_debug_jacobi(dbg, 0)
nb = max(dbg[4], 1)
print(f"ms per step {dt * 1e3:.1f} | blocks with >= 64 vectors per step {dbg[4] / K:.0f}, sweeps/blk {dbg[3] / nb:.2f}, QR kclk/blk {dbg[5] / nb / 1e3:.0f}, "
      f"Jacobi kclk/blk {dbg[6] / nb / 1e3:.0f}")
