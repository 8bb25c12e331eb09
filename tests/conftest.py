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
