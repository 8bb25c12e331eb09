import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

GOLDEN = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


def has_gpu():
    try:
        import torch
        return torch.cuda.is_available()
    except Exception:
        return False


def pytest_collection_modifyitems(config, items):
    if has_gpu():
        return
    skip = pytest.mark.skip(reason="no CUDA device")
    for it in items:
        if "gpu" in it.keywords:
            it.add_marker(skip)


# ---- conversions between the oracle's MPS and the product package's host container ----
def to_host(m):
    import optimalcontrolmps_b200 as oc
    return oc.IQMPS(m.A, m.q, m.llim, m.rlim)


def to_oracle(h):
    from oracle import bh_mps as ob
    return ob.MPS(h.A, [np.asarray(x, dtype=np.int64) for x in h.q], h.llim, h.rlim)


def golden_state(z, prefix):
    from oracle import bh_mps as ob
    L = sum(1 for k in z.files if k.startswith(prefix + "_A"))
    return ob.MPS([z[f"{prefix}_A{j}"] for j in range(L)], [z[f"{prefix}_q{b}"].astype(np.int64) for b in range(L + 1)])


def load_golden(name):
    return np.load(os.path.join(GOLDEN, name))


def random_symmetric_mps(L, D, Npart, chi, seed):
    """Random number-conserving MPS with bond dimensions up to ``chi`` (flat spectra: a stress input), canonical
    with the centre at site 1.  Built with the oracle's containers."""
    from oracle import bh_mps as ob
    rng = np.random.default_rng(seed)
    # number of left-half configurations per charge bounds the multiplicity of each label
    ways = [np.zeros(Npart + 1, dtype=object) for _ in range(L + 1)]
    ways[0][0] = 1
    for b in range(1, L + 1):
        for q in range(Npart + 1):
            ways[b][q] = sum(ways[b - 1][q - s] for s in range(D) if 0 <= q - s)
    waysR = [np.zeros(Npart + 1, dtype=object) for _ in range(L + 1)]
    waysR[L][Npart] = 1
    for b in range(L - 1, -1, -1):
        for q in range(Npart + 1):
            waysR[b][q] = sum(waysR[b + 1][q + s] for s in range(D) if q + s <= Npart)
    qs = []
    for b in range(L + 1):
        cap = np.array([min(int(ways[b][q]), int(waysR[b][q])) for q in range(Npart + 1)])
        allowed = np.nonzero(cap > 0)[0]
        mult = np.zeros(Npart + 1, dtype=int)
        want = min(chi, int(cap.sum()))
        while mult.sum() < want:                     # round-robin fill, centre charges first
            order = sorted(allowed, key=lambda q: abs(q - Npart * b / L))
            for q in order:
                if mult.sum() < want and mult[q] < cap[q]:
                    mult[q] += 1
        qs.append(np.repeat(np.arange(Npart + 1), mult).astype(np.int64))
    A = []
    s = np.arange(D)
    for j in range(L):
        ql, qr = qs[j], qs[j + 1]
        a = rng.normal(size=(len(ql), D, len(qr))) + 1j * rng.normal(size=(len(ql), D, len(qr)))
        ok = (ql[:, None, None] + s[None, :, None]) == qr[None, None, :]
        A.append(a * ok)
    psi = ob.MPS(A, qs, llim=0, rlim=L + 1)
    psi.position(1)
    psi.normalize()
    return psi
