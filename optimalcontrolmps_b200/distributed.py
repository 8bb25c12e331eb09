"""Multi-GPU plumbing for the parts of the path that shard naturally.

* GRAPE/GROUP Hessian: rows r = 1..Nt-2 are independent given (psi_t, xiHlist, divT) and write disjoint
  entries (reference src/OptimalControl.cpp:252-279, work queue :305-335).  Rows are dealt to the ranks so
  that every rank propagates about the same number of Trotter steps; each rank computes its rows on its
  own GPU and the row blocks are combined with ONE all-gather (NCCL over NVLink; gloo in the CPU tests).
* Batches of independent controls (seeds): replicas, results combined with one all-gather.

A single cost+gradient evaluation never leaves one GPU.  torch.distributed is used only as plumbing.
"""
from __future__ import annotations

from typing import List, Sequence

import numpy as np


def partition_rows(Nt: int, world: int, rank: int) -> List[int]:
    """Rows of the Hessian owned by ``rank``.  Row r costs (Nt-2-r) Trotter steps, so rows are dealt in a
    zig-zag (boustrophedon) order over the cost-sorted list, which balances the step counts to within one row."""
    rows = list(range(1, Nt - 1))            # already sorted by decreasing cost
    mine = []
    for i, r in enumerate(rows):
        lap, pos = divmod(i, world)
        owner = pos if lap % 2 == 0 else world - 1 - pos
        if owner == rank:
            mine.append(r)
    return mine


def row_cost(Nt: int, rows: Sequence[int]) -> int:
    return sum(Nt - 2 - r + 1 for r in rows)


def pack_rows(Hf: np.ndarray, rows: Sequence[int], max_rows: int) -> np.ndarray:
    """Row block of this rank: [row index, H[r, r:]] per owned row, padded to max_rows rows."""
    N = Hf.shape[0]
    buf = np.full((max_rows, N + 1), -1.0)
    for k, r in enumerate(rows):
        buf[k, 0] = r
        buf[k, 1:] = 0.0
        buf[k, 1 + r:] = Hf[r, r:]
    return buf


def unpack_rows(blocks: Sequence[np.ndarray], N: int) -> np.ndarray:
    """Inverse of pack_rows over all ranks: the symmetric fidelity part of the Hessian."""
    H = np.zeros((N, N))
    for buf in blocks:
        for line in buf:
            r = int(round(line[0]))
            if r < 0:
                continue
            H[r, r:] = line[1 + r:]
            H[r:, r] = line[1 + r:]
    return H


def allgather_array(local: np.ndarray, device=None) -> List[np.ndarray]:
    """One all-gather of equally shaped float64 arrays over the default process group."""
    import torch
    import torch.distributed as dist
    world = dist.get_world_size()
    t = torch.from_numpy(np.ascontiguousarray(local, dtype=np.float64))
    if device is not None:
        t = t.to(device)
    outs = [torch.empty_like(t) for _ in range(world)]
    dist.all_gather(outs, t)
    return [o.cpu().numpy() for o in outs]


def sharded_hessian(oc, control, new_control: bool = True, device=None, convert: bool = False) -> np.ndarray:
    """Hessian with rows sharded over the ranks of the default process group (``calcHessian_parallel`` with the row
    work queue of src/OptimalControl.cpp:305-335 dealt to GPUs instead of threads).  ``oc`` is an
    optimalcontrolmps_b200.OptimalControl living on this rank's GPU.  Every rank runs the prerequisites (both sweeps,
    K.xi) itself -- its rows trail its own psi sweep, so they are not on the critical path -- and its share of the
    rows; the row blocks are combined with ONE all-gather.  Returns the GRAPE Hessian in u (N x N), or with
    ``convert=True`` what ``getHessian`` returns (GROUP: V^T H V, src/ControlBasis.cpp:95-119)."""
    import torch.distributed as dist
    world, rank = dist.get_world_size(), dist.get_rank()
    N = oc.getN()
    oc.rows = partition_rows(N, world, rank)
    try:
        u = control if oc.GRAPE else oc.basis.convertControl(control, new_control)
        oc._calcHessian(u, new_control)
        max_rows = max(len(partition_rows(N, world, r)) for r in range(world))
        blocks = allgather_array(pack_rows(oc._hessian_fidelity_part, oc.rows, max_rows), device)
    finally:
        oc.rows = None
    H = oc._hessian_reg_part + unpack_rows(blocks, N)
    if convert and not oc.GRAPE:
        return np.array(oc.basis.convertHessian(H))
    return H
