"""CPU tests of the product package: host-side mirror of the reference interface (no compute),
the C-ABI library loads and exports every symbol include/ocmps.h declares, and it refuses to run
without a GPU (no CPU fallback)."""
import ctypes
import os
import re
import subprocess
import sys

import numpy as np
import pytest

import optimalcontrolmps_b200 as oc
from optimalcontrolmps_b200 import _lib, distributed
from conftest import ROOT, has_gpu, load_golden
from test_oracle_goldens import check_control_basis


def test_control_basis_matches_reference_goldens():           # tests/ControlBasisTests.cpp through the product classes
    N, M = 5, 4
    simple = oc.ControlBasis([1.0] * N, [1.0] * N, [[2.0] * M for _ in range(N)])
    u0 = [1, 1.1, 1.2, 1.3, 1.4, 1.5, 1.6, 1.7, 1.8, 1.9, 2]
    chopped = oc.ControlBasisFactory.buildChoppedSineBasis(u0, 1e-1, 1.0, 5)
    check_control_basis(simple, chopped)


def test_seed_generator():
    x = oc.SeedGenerator.linspace(2.0, 50.0, 11)               # include/SeedGenerator.hpp:26-37
    assert len(x) == 11 and x[0] == 2.0 and abs(x[-1] - 50.0) < 1e-9
    assert len(oc.SeedGenerator.generateRange(0.0, 0.5, 2.0)) == 5
    u = oc.SeedGenerator.linsigmoidSeed(2.5, 50, 201, np.random.default_rng(1))
    from oracle import optimal_control as oo
    r = np.random.default_rng(1)
    a_, c_, d_ = float(r.uniform(0.01, 0.15)), float(r.uniform(0.06, 0.18)), float(r.uniform(60, 80))
    assert np.allclose(u, oo.linsigmoid_seed(2.5, 50, 201, a_, c_, d_))            # include/SeedGenerator.hpp:66-95
    assert len(u) == 201 and abs(u[0] - 2.5) < 0.02 and abs(u[-1] - 50) < 0.5
    assert all(2.0 <= v <= 100.0 for v in u)                   # src/BH_nlp.cpp:55-56 bounds
    a = oc.SeedGenerator.adiabaticSeed(2.5, 50, 101)
    assert len(a) == 101 and abs(a[-1] - 50) < 1e-9
    assert np.allclose(a, oo.adiabatic_seed(2.5, 50, 101))
    c = oc.SeedGenerator.randomCoeffSeed(-4, 4, 10, np.random.default_rng(2))
    assert len(c) == 10 and all(-4 <= v <= 4 for v in c)


def test_args_and_containers():
    a = oc.Args("Cutoff=", 1e-8, "Maxm=", 100)
    assert a.defined("Cutoff") and a.defined("Maxm=") and a.getInt("Maxm") == 100 and a.getReal("Cutoff=") == 1e-8
    assert not oc.Args("Cutoff", 1e-7).defined("Maxm")
    z = load_golden("golden_L5.npz")
    A = [z[f"init_A{j}"] for j in range(5)]
    q = [z[f"init_q{b}"] for b in range(6)]
    psi = oc.IQMPS(A, q)
    assert psi.N() == 5 and psi.D == 6 and psi.bond_dims()[0] == 1
    with pytest.raises(ValueError):
        oc.IQMPS(A, q[:-1])
    s = oc.BoseHubbard(20, 5)
    assert s.N() == 20 and s.D == 6


def test_ground_state_fixtures_load():
    from optimalcontrolmps_b200.states import ground_state
    psi = ground_state(20, 5, 20, 2.5)
    assert psi.N() == 20 and psi.D == 6 and max(psi.bond_dims()) <= 100
    assert ground_state(8, 4, 8, 50).N() == 8


def declared_symbols():
    text = open(os.path.join(ROOT, "include", "ocmps.h")).read()
    return sorted(set(re.findall(r"\b(ocmps_[a-z_A-Z0-9]+)\s*\(", text)))


def test_library_exports_every_declared_symbol():
    assert os.path.exists(_lib.LIB_PATH), "libocmps.so missing: run `make` (or __graft_entry__.build())"
    lib = ctypes.CDLL(_lib.LIB_PATH)
    names = declared_symbols()
    assert len(names) >= 30
    for n in names:
        assert hasattr(lib, n), f"{n} declared in include/ocmps.h but not exported"
    assert sorted(_lib.SYMBOLS) == names                        # the ctypes table covers the whole header
    assert _lib.load().ocmps_version() >= 100


@pytest.mark.skipif(has_gpu(), reason="checks the behaviour on a machine without a GPU")
def test_fails_loudly_without_gpu():
    with pytest.raises(_lib.OcmpsError) as e:
        oc.Context(0)
    assert "no CPU fallback" in str(e.value) or "CUDA" in str(e.value)


def test_row_partition_is_balanced_and_complete():
    for Nt, world in [(201, 8), (201, 2), (11, 4), (5, 8)]:
        allrows = []
        costs = []
        for r in range(world):
            rows = distributed.partition_rows(Nt, world, r)
            allrows += rows
            costs.append(distributed.row_cost(Nt, rows))
        assert sorted(allrows) == list(range(1, Nt - 1))
        if Nt > 4 * world:
            assert max(costs) - min(costs) <= Nt            # within one row's cost
    H = np.random.default_rng(0).normal(size=(9, 9))
    H = H + H.T
    H[0] = 0; H[:, 0] = 0; H[8] = 0; H[:, 8] = 0
    blocks = [distributed.pack_rows(H, distributed.partition_rows(9, 3, r), 3) for r in range(3)]
    assert np.allclose(distributed.unpack_rows(blocks, 9), H)


WORKER = r'''
import os, sys
sys.path.insert(0, sys.argv[1])
import numpy as np, torch, torch.distributed as dist
from optimalcontrolmps_b200 import distributed as d
dist.init_process_group("gloo", init_method="tcp://127.0.0.1:%s" % sys.argv[2], rank=int(sys.argv[3]), world_size=2)
N = 13
rng = np.random.default_rng(42)
H = rng.normal(size=(N, N)); H = H + H.T
H[0] = 0; H[:, 0] = 0; H[N - 1] = 0; H[:, N - 1] = 0
rank = dist.get_rank()
rows = d.partition_rows(N, 2, rank)
local = np.zeros((N, N))
for r in rows:
    local[r, r:] = H[r, r:]; local[r:, r] = H[r:, r]
blocks = d.allgather_array(d.pack_rows(local, rows, max(len(d.partition_rows(N, 2, k)) for k in range(2))))
full = d.unpack_rows(blocks, N)
assert np.allclose(full, H), "gathered Hessian differs"
print("rank", rank, "ok")
dist.destroy_process_group()
'''


def test_gloo_world_size_2_row_gather(tmp_path):
    script = tmp_path / "worker.py"
    script.write_text(WORKER)
    port = str(29500 + os.getpid() % 2000)
    procs = [subprocess.Popen([sys.executable, str(script), ROOT, port, str(r)], stdout=subprocess.PIPE, stderr=subprocess.STDOUT)
             for r in range(2)]
    outs = [p.communicate(timeout=180)[0].decode() for p in procs]
    for p, o in zip(procs, outs):
        assert p.returncode == 0, o
        assert "ok" in o


def test_site_operator_diagonals_and_queue_default():
    """Host-side pieces of the observables: the operator diagonals are those of include/BH_sites.h:129-171 (N: n,
    N(N-1): n^2 - n, NN: n^2), and importing the package asks for one hardware queue per chain stream unless the user
    has chosen a value."""
    import os
    import numpy as np
    import optimalcontrolmps_b200 as oc
    n = np.arange(6.0)
    ops = oc.SliceStore.SITE_OPS
    assert np.array_equal(ops["N"](n), n)
    assert np.array_equal(ops["N(N-1)"](n), n * n - n)
    assert np.array_equal(ops["NN"](n), n * n)
    assert np.array_equal(ops["Id"](n), np.ones(6))
    assert os.environ.get("CUDA_DEVICE_MAX_CONNECTIONS") is not None
