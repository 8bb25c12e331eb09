"""What ONE rank of a `world`-GPU sharded cfg3 Hessian does, timed on one GPU: both sweeps, K.xi and the rows of rank `rank`
as one schedule (ocmps_hessian_eval through OptimalControl._calcHessian).  usage: gpu_hess_rank_share.py world [rank] [reps]
Set OCMPS_HESSIAN_TRACE=1 for the device-side timeline."""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import bench
import optimalcontrolmps_b200 as oc
from optimalcontrolmps_b200 import distributed as ocd
from optimalcontrolmps_b200.states import ground_state

CFG = bench.CFG
world = int(sys.argv[1]) if len(sys.argv) > 1 else 8
rank = int(sys.argv[2]) if len(sys.argv) > 2 else 0
reps = int(sys.argv[3]) if len(sys.argv) > 3 else 3
L, d = CFG["L"], CFG["d"]
st = oc.BH_tDMRG(oc.BoseHubbard(L, d), CFG["J"], CFG["tstep"], oc.Args("Cutoff=", CFG["cutoff"], "Maxm=", CFG["maxm"]))
psi_i, psi_f = ground_state(L, d, CFG["Npart"], CFG["U_i"]), ground_state(L, d, CFG["Npart"], CFG["U_f"])
basis, c = bench.make_hessian_problem_host(0)
o = oc.OptimalControl(psi_f, psi_i, st, basis, CFG["gamma"])
o.setThreadCount(4)
N = o.getN()
o.rows = ocd.partition_rows(N, world, rank) if world > 1 else None
u = o.basis.convertControl(list(c), True)
print("world", world, "rank", rank, "rows", len(o.rows) if o.rows else N - 2, "row steps", ocd.row_cost(N, o.rows) if o.rows else (N - 2) * (N - 1) // 2)
for r in range(reps):
    t0 = time.perf_counter()
    o._calcHessian(u, True)
    print("call", r, "wall s", time.perf_counter() - t0, flush=True)
