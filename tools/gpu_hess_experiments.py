"""Timing experiments on one rank's share of the 8-GPU cfg3 Hessian (one GPU): row subsets, rows in flight, overlap share.
usage: gpu_hess_experiments.py [full]   ('full' adds the complete 1-GPU Hessian with and without the chunk overlaps)"""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import bench
import optimalcontrolmps_b200 as oc
from optimalcontrolmps_b200 import distributed as ocd
from optimalcontrolmps_b200.states import ground_state

CFG = bench.CFG
L, d = CFG["L"], CFG["d"]
st = oc.BH_tDMRG(oc.BoseHubbard(L, d), CFG["J"], CFG["tstep"], oc.Args("Cutoff=", CFG["cutoff"], "Maxm=", CFG["maxm"]))
psi_i, psi_f = ground_state(L, d, CFG["Npart"], CFG["U_i"]), ground_state(L, d, CFG["Npart"], CFG["U_f"])
basis, c = bench.make_hessian_problem_host(0)
o = oc.OptimalControl(psi_f, psi_i, st, basis, CFG["gamma"])
o.setThreadCount(4)
N = o.getN()
u = o.basis.convertControl(list(c), True)


def run(tag, rows, reps=2, chains=None, env=None):
    o.rows = rows
    o.hessian_chains = chains
    for k, v in (env or {}).items():
        os.environ[k] = v
    ts = []
    for r in range(reps):
        t0 = time.perf_counter()
        o._calcHessian(u, True)
        ts.append(time.perf_counter() - t0)
    for k in (env or {}):
        os.environ.pop(k, None)
    print(f"{tag}: rows {len(rows) if rows is not None else N - 2} steps {ocd.row_cost(N, rows) if rows is not None else 0} "
          f"wall {' '.join(f'{t:.3f}' for t in ts)}", flush=True)


z0, z7 = ocd.partition_rows(N, 8, 0), ocd.partition_rows(N, 8, 7)
run("warm-up zigzag rank 0", z0, reps=2)
os.environ["OCMPS_HESSIAN_TRACE"] = "1"
run("zigzag rank 0", z0)
run("zigzag rank 7", z7)
run("prerequisites only (one short row)", [N - 2])
run("zigzag rank 0, no chunk overlaps", z0, env={"OCMPS_HESSIAN_SKIP_OVL": "1"})
run("zigzag rank 0, 12 rows in flight", z0, chains=12)
run("zigzag rank 0, 25 rows in flight", z0, chains=25)
run("contiguous rows 1..19", list(range(1, 20)))
run("contiguous rows 153..199", list(range(153, 200)))
run("contiguous rows 102..125", list(range(102, 126)))
if len(sys.argv) > 1 and sys.argv[1] == "full":
    run("1 GPU, all rows", None, reps=2)
    run("1 GPU, all rows, no chunk overlaps", None, reps=1, env={"OCMPS_HESSIAN_SKIP_OVL": "1"})
