"""Worker of tests/test_gpu_sharded_hessian.py (launched with torch.distributed.run, 2 ranks): the Hessian with rows sharded
over the ranks (NCCL when every rank has its own GPU, gloo when they share one) against the same Hessian computed by rank 0
alone, and against the golden (oracle) Hessian."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))


def main():
    import torch
    import torch.distributed as dist
    from conftest import golden_state, load_golden, to_host
    rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
    ndev = torch.cuda.device_count()
    own_gpu = ndev >= world
    devi = rank if own_gpu else 0
    torch.cuda.set_device(devi)
    dev = torch.device("cuda", devi)
    if own_gpu:
        dist.init_process_group("nccl", device_id=dev)
    else:
        dist.init_process_group("gloo")
    import optimalcontrolmps_b200 as oc
    from optimalcontrolmps_b200 import distributed as ocd
    z = load_golden(sys.argv[1])
    L, d, Np, J, cs, ce, T, ts, cutoff, maxm, M, gamma, N = z["params"]
    L, d, N, M = int(L), int(d), int(N), int(M)
    init, target = golden_state(z, "init"), golden_state(z, "target")
    cap = max([int(maxm)] + init.bond_dims() + target.bond_dims())
    ctx = oc.Context.default(devi)
    st = oc.BH_tDMRG(oc.BoseHubbard(L, d), J, ts, oc.Args("Cutoff=", cutoff, "Maxm=", int(maxm)), chi_cap=cap, ctx=ctx)
    u = list(z["u"])
    ocg = oc.OptimalControl(to_host(target), to_host(init), st, N, float(gamma))
    H = ocd.sharded_hessian(ocg, u, True, dev if own_gpu else None)
    # GROUP flavour through the basis
    u0 = oc.SeedGenerator.linspace(float(cs), float(ce), N)
    basis = oc.ControlBasisFactory.buildChoppedSineBasis(u0, float(ts), float(T), M)
    ocb = oc.OptimalControl(to_host(target), to_host(init), st, basis, float(gamma))
    c = list(z["c"])
    Hg = ocd.sharded_hessian(ocb, c, True, dev if own_gpu else None, convert=True)
    if rank == 0:
        H1 = np.array(ocg.getHessian(u, True))
        Hg1 = np.array(ocb.getHessian(c, True))
        scale = np.max(np.abs(H1))
        print("RESULT", float(np.max(np.abs(H - H1)) / scale), float(np.max(np.abs(Hg - Hg1)) / np.max(np.abs(Hg1))),
              float(np.max(np.abs(H - z["hessian"])) / np.max(np.abs(z["hessian"]))),
              float(np.max(np.abs(Hg - z["group_hessian"])) / np.max(np.abs(z["group_hessian"]))), "nccl" if own_gpu else "gloo", flush=True)
    dist.barrier()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
