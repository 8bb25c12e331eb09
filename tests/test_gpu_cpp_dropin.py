"""Drives the C++ drop-in layer (cpp/: the reference's class API over libocmps) on the GPU: the problem and the
oracle's expected results are handed to cpp/tests/test_dropin as a binary file; the C++ program uses
BoseHubbard / BH_tDMRG / OptimalControl / ControlBasisFactory / BH_nlp exactly like the reference's own callers."""
import os
import struct
import subprocess

import numpy as np
import pytest

from conftest import ROOT, golden_state, load_golden

pytestmark = pytest.mark.gpu

BIN = os.path.join(ROOT, "cpp", "tests", "test_dropin")


def w_vec(f, a):
    a = np.ascontiguousarray(a, dtype=np.float64).ravel()
    f.write(struct.pack("<i", a.size))
    f.write(a.tobytes())


def w_ivec(f, a):
    a = np.ascontiguousarray(a, dtype=np.int32).ravel()
    f.write(struct.pack("<i", a.size))
    f.write(a.tobytes())


def w_state(f, psi):
    w_ivec(f, psi.bond_dims())
    w_ivec(f, np.concatenate(psi.q))
    w_vec(f, np.concatenate([a.ravel() for a in psi.A]).view(np.float64))


@pytest.mark.parametrize("name", ["golden_L5.npz", "golden_L6_maxm.npz"])
def test_cpp_dropin(name, tmp_path):
    if not os.path.exists(BIN):
        subprocess.run(["make", "-C", os.path.join(ROOT, "cpp")], check=True)
    z = load_golden(name)
    L, d, Np, J, cs, ce, T, ts, cutoff, maxm, M, gamma, N = z["params"]
    init, target = golden_state(z, "init"), golden_state(z, "target")
    maxm = int(maxm)
    D = int(d) + 1
    cap = max([maxm] + init.bond_dims() + target.bond_dims()) if maxm > 0 else min(D ** (int(L) // 2), 256)
    path = tmp_path / "problem.bin"
    with open(path, "wb") as f:
        f.write(struct.pack("<6i", int(L), int(d), int(N), int(M), maxm if maxm > 0 else 0, int(cap)))
        f.write(struct.pack("<7d", J, ts, T, cutoff, gamma, cs, ce))
        w_state(f, init)
        w_state(f, target)
        w_vec(f, z["u"]); w_vec(f, z["c"])
        f.write(struct.pack("<d", float(z["cost"])))
        w_vec(f, z["fidelities"]); w_vec(f, z["grad"]); w_vec(f, z["hessian"])
        f.write(struct.pack("<d", float(z["group_cost"])))
        w_vec(f, z["group_grad"]); w_vec(f, z["group_hessian"])
    res = subprocess.run([BIN, str(path)], capture_output=True, text=True, timeout=600, cwd=str(tmp_path))
    print(res.stdout)
    print(res.stderr)
    assert res.returncode == 0, res.stdout + res.stderr
    assert "PASSED" in res.stdout
