"""Per-step wall time of the cfg2 forward (psi) and backward (xi) sweeps on bench.py's control (GPU).
Usage: python tools/gpu_sweep_profile.py [stride]   -- prints every `stride`-th step: index, ms, max bond dimension."""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import bench
import optimalcontrolmps_b200 as oc
from optimalcontrolmps_b200.states import ground_state

CFG = bench.CFG
stride = int(sys.argv[1]) if len(sys.argv) > 1 else 10
L, d = CFG["L"], CFG["d"]
st = oc.BH_tDMRG(oc.BoseHubbard(L, d), CFG["J"], CFG["tstep"], oc.Args("Cutoff=", CFG["cutoff"], "Maxm=", CFG["maxm"]))
basis, c, u = bench.make_problem_host(0)
N = len(u)
for name, start, fwd in (("psi", ground_state(L, d, CFG["Npart"], CFG["U_i"]), True),):
    for rep in range(2):          # second pass: step graphs are captured
        pd = st.to_device(start)
        ts = []
        dims = []
        for k in range(N - 1):
            i = k if fwd else N - 1 - k
            t0 = time.perf_counter()
            st.step(pd, u[i], u[i + 1] if fwd else u[i - 1], fwd)
            ts.append(time.perf_counter() - t0)
            if k % stride == 0 or k == N - 2:
                dims.append((k, max(pd.bond_dims())))
    ts = np.array(ts) * 1e3
    print(name, "total ms", ts.sum())
    for (k, m) in dims:
        print(f"  {name} step {k:3d}  {ts[k]:7.3f} ms  chi_max {m}")
