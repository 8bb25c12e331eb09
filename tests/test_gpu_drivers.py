"""The C++ drivers over the drop-in classes: cpp/main/OptimizeRamp (reference main/OptimizeRamp.cpp: InputGroup
file -> InitializeState -> GROUP problem -> TNLP adapter driven by the optimiser behind the IpoptApplication calls -> output files) and
cpp/main/AmoebaOpt (reference main/AmoebaOpt.cpp: Nelder-Mead with the bound penalty, independent evaluations on concurrent problem
copies).  Small problem (L=5), a few iterations; checks the contract of the files and that the cost goes down."""
import os
import subprocess

import numpy as np
import pytest

pytestmark = pytest.mark.gpu

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

INPUT = """
input
{
    N = 5
    Npart = 5
    d = 4
    T = 0.1            # 11 time points
    tstep = 0.01
    M = 3
    gamma = 1e-6
    maxBondDim = 20
    threshold = 1e-8
    cacheProgress = yes
    useBFGS = %s
    maxIter = %d
    optTol = 1e-9
    threadCount = 2
    maxFun = 40
    parallelEvals = 3
}
"""


def _build():
    r = subprocess.run(["make", "-C", os.path.join(ROOT, "cpp")], capture_output=True, text=True)
    assert r.returncode == 0, r.stderr[-2000:]


def _run(tmp_path, exe, bfgs, iters):
    _build()
    inp = tmp_path / "InputFile_BHcontrol"
    inp.write_text(INPUT % ("yes" if bfgs else "no", iters))
    p = subprocess.run([os.path.join(ROOT, "cpp", "main", exe), str(inp), "3"], cwd=tmp_path, capture_output=True, text=True, timeout=900)
    assert p.returncode == 0, p.stdout[-3000:] + p.stderr[-3000:]
    return p.stdout


@pytest.mark.parametrize("bfgs", [False, True])
def test_optimize_ramp_driver(tmp_path, bfgs):
    out = _run(tmp_path, "OptimizeRamp", bfgs, 3)
    fs = [float(l.split("f =")[1].split()[0]) for l in out.splitlines() if l.startswith("iter ") and "f =" in l]
    assert len(fs) >= 2 and all(b <= a + 1e-14 for a, b in zip(fs, fs[1:])) and fs[-1] < fs[0]          # monotone descent
    ramp = np.loadtxt(tmp_path / "BHrampInitialFinal.txt")
    assert ramp.shape == (11, 5)                                       # time, u_initial, F_initial, u_final, F_final
    assert np.all(ramp[:, 1] >= 2.0) and np.all(ramp[:, 3] <= 100.0) and np.all((ramp[:, [2, 4]] >= -1e-12) & (ramp[:, [2, 4]] <= 1 + 1e-12))
    assert ramp[-1, 4] >= ramp[-1, 2] - 1e-12                          # the optimised ramp ends with at least the initial fidelity
    expn = np.loadtxt(tmp_path / "ExpectationN.txt")
    assert expn.shape == (11, 6) and np.allclose(expn[:, 1:].sum(axis=1), 5.0, atol=1e-9)                 # <N_j> per slice
    cache = np.loadtxt(tmp_path / "ProgressCache.txt", ndmin=2)
    assert cache.shape[1] == 4 and cache.shape[0] == len(fs)
    Hgroup = np.loadtxt(tmp_path / "GROUPHessian.txt")
    Hgrape = np.loadtxt(tmp_path / "GRAPEHessian.txt")
    assert Hgroup.shape == (3, 3) and Hgrape.shape == (11, 11)
    assert np.allclose(Hgroup, Hgroup.T, atol=1e-12) and np.allclose(Hgrape, Hgrape.T, atol=1e-12)


def test_amoeba_driver(tmp_path):
    out = _run(tmp_path, "AmoebaOpt", False, 3)
    hist = np.loadtxt(tmp_path / "AmoebaHistory.txt", ndmin=2)
    assert hist[0, 1] >= hist[-1, 1] and hist[-1, 2] >= 4              # best cost never increases; the simplex was built (M+1 evaluations)
    res = open(tmp_path / "AmoebaResult.txt").read().split()
    assert len(res) == 1 + 3 and float(res[0]) <= hist[0, 1] + 1e-14
    ramp = np.loadtxt(tmp_path / "BHrampInitialFinal.txt")
    assert ramp.shape == (11, 5)
    assert "Initialize" in out
