"""Time the GRAPE Hessian (cfg3 shape, Nt points) for several numbers of rows in flight on one GPU."""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import bench

def main():
    Nt = int(sys.argv[1]) if len(sys.argv) > 1 else 201
    chains = [int(x) for x in sys.argv[2:]] or [16, 32]
    import torch
    import optimalcontrolmps_b200 as oc
    from optimalcontrolmps_b200.states import ground_state
    CFG = bench.CFG
    ctx = oc.Context.default(0)
    st = oc.BH_tDMRG(oc.BoseHubbard(CFG["L"], CFG["d"]), CFG["J"], CFG["tstep"], oc.Args("Cutoff=", CFG["cutoff"], "Maxm=", CFG["maxm"]), ctx=ctx)
    psi_i = ground_state(CFG["L"], CFG["d"], CFG["Npart"], CFG["U_i"])
    psi_f = ground_state(CFG["L"], CFG["d"], CFG["Npart"], CFG["U_f"])
    u = list(np.linspace(bench.CFG["U_i"], 30.0, Nt))
    for nch in chains:
        och = oc.OptimalControl(psi_f, psi_i, st, Nt, bench.CFG["gamma"])
        och.setThreadCount(4)
        och.hessian_chains = nch
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        H = np.array(och.getHessian(u, True))
        torch.cuda.synchronize()
        print("chains", nch, "wall_s", time.perf_counter() - t0, "checksum", float(np.abs(H).sum()), flush=True)

if __name__ == "__main__":
    main()
