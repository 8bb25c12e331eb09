"""The engine's expert switches select alternative kernel paths for the same arithmetic (dense-GEMM fall-backs of the
sector-wise merge and gauge push, plain launches instead of CUDA graphs, size of the batched overlap passes of a Hessian
row).  Every path must give the same cost / gradient / Hessian and the same bond dimensions; paths that only regroup
launches must be bit-identical."""
import json
import os
import subprocess
import sys

import numpy as np
import pytest

pytestmark = pytest.mark.gpu

HERE = os.path.dirname(os.path.abspath(__file__))


def probe(golden, **env):
    e = dict(os.environ)
    e.update({k: str(v) for k, v in env.items()})
    out = subprocess.run([sys.executable, os.path.join(HERE, "_switch_probe.py"), golden], env=e, capture_output=True, text=True,
                         timeout=600)
    assert out.returncode == 0, out.stderr[-2000:]
    line = [ln for ln in out.stdout.splitlines() if ln.startswith("PROBE ")][-1]
    d = json.loads(line[6:])
    return d["cost"], np.array(d["grad"]), np.array(d["hess"]), d["dims"]


@pytest.fixture(scope="module")
def default_run():
    return probe("golden_L8_maxm.npz")


@pytest.mark.parametrize("env,exact", [
    ({"OCMPS_GRAPH": 0}, True),                  # same kernels, launched one by one
    ({"OCMPS_HESSIAN_CHUNK": 3}, True),          # same overlaps, batched in passes of 3 slices instead of 32
    ({"OCMPS_FUSED_MERGE": 0}, False),           # dense DMMA GEMM + gate kernel instead of the sector-wise merge+gate
    ({"OCMPS_FUSED_PUSH": 0}, False),            # dense GEMM after a gauge move instead of the push inside build_factors
    ({"OCMPS_TMA_COPY": 1}, True),               # slice store through TMA bulk copies instead of loads and stores: the same bytes
    ({"OCMPS_GRAM": 1}, False),                  # Gram matrix (DMMA) + pivoted Cholesky instead of the Householder QR in the gate decompositions
])
def test_switch_gives_the_same_result(default_run, env, exact):
    c0, g0, h0, d0 = default_run
    c1, g1, h1, d1 = probe("golden_L8_maxm.npz", **env)
    assert d1 == d0
    if exact:
        assert c1 == c0 and np.array_equal(g1, g0) and np.array_equal(h1, h0)
    else:
        assert abs(c1 - c0) <= 1e-12 * abs(c0)
        assert np.max(np.abs(g1 - g0)) <= 1e-11 * np.max(np.abs(g0))
        assert np.max(np.abs(h1 - h0)) <= 1e-10 * np.max(np.abs(h0))


def test_gram_path_on_the_full_size_golden():
    """OCMPS_GRAM=1 only takes charge blocks with at least 48 vectors, which the small goldens above do not have: the complete
    cfg2 evaluation (every bond dimension of 2 x 201 slices, cost 1e-9, gradient 1e-7) is repeated with the Gram-matrix (DMMA) +
    pivoted-Cholesky path in a process of its own."""
    e = dict(os.environ)
    e["OCMPS_GRAM"] = "1"
    out = subprocess.run([sys.executable, "-m", "pytest", "-x", "-q", "-m", "gpu",
                          os.path.join(HERE, "test_gpu_fullsize.py::test_cfg2_full_evaluation_matches_golden")],
                         env=e, capture_output=True, text=True, timeout=900, cwd=os.path.dirname(HERE))
    assert out.returncode == 0, out.stdout[-2000:] + out.stderr[-2000:]
