"""Slice-store copies (work MPS <-> store slot) at the cfg2 shape with every bulk bond at chi=100: device time and achieved
bandwidth of a put + get pair; checks that the slice comes back bit for bit.  OCMPS_TMA_COPY=0 selects the plain load/store kernels."""
import sys, os, ctypes, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import numpy as np
import optimalcontrolmps_b200 as oc
from optimalcontrolmps_b200 import _lib
from conftest import random_symmetric_mps, to_host

L, D, Np, chi = 20, 6, 20, 100
psi = random_symmetric_mps(L, D, Np, chi, seed=3)
st = oc.BH_tDMRG(oc.BoseHubbard(L, D - 1), 1.0, 1e-2, oc.Args("Cutoff=", 1e-8, "Maxm=", chi))
dev = st.to_device(to_host(psi))
store = st.new_store(40)
host0 = dev.download()
nbytes = sum(a.size for a in host0.A) * 16
lib = st.ctx.lib
for s in range(40):
    store.put(s, dev)
back = store.get(17).download()
same = all(np.array_equal(a, b) for a, b in zip(host0.A, back.A)) and all(np.array_equal(a, b) for a, b in zip(host0.q, back.q))
print("slice bytes", nbytes, "round trip bit-identical:", same)
reps = 200
ms = ctypes.c_double()
_lib.check(lib.ocmps_timer_start(st.ctx.h))
for r in range(reps):
    store.put(r % 40, dev)
_lib.check(lib.ocmps_timer_stop(st.ctx.h, ctypes.byref(ms)))
t_put = ms.value / reps
tmp = store.get(0)
_lib.check(lib.ocmps_timer_start(st.ctx.h))
for r in range(reps):
    lib.ocmps_store_get(store.h, r % 40, tmp.h)
_lib.check(lib.ocmps_timer_stop(st.ctx.h, ctypes.byref(ms)))
t_get = ms.value / reps
print(f"put {t_put * 1e3:.2f} us ({2 * nbytes / t_put / 1e9:.2f} TB/s read+write incl. call overhead), get {t_get * 1e3:.2f} us ({2 * nbytes / t_get / 1e9:.2f} TB/s)")
assert same
