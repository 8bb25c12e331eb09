"""Loading / saving of charge-labelled MPS fixtures (flat complex128 tensors + bond charges)."""
from __future__ import annotations

import os
import numpy as np

from .api import IQMPS

DATA_DIR = os.path.join(os.path.dirname(os.path.abspath(__file__)), "data")


def load_state(path: str) -> IQMPS:
    z = np.load(path)
    L = sum(1 for k in z.files if k.startswith("A"))
    return IQMPS([z[f"A{j}"] for j in range(L)], [z[f"q{b}"] for b in range(L + 1)], 0, 2)


def save_state(path: str, psi: IQMPS):
    arrs = {f"A{j}": a for j, a in enumerate(psi.A)}
    arrs.update({f"q{b}": x for b, x in enumerate(psi.q)})
    np.savez_compressed(path, **arrs)


def ground_state(L: int, d: int, Npart: int, U: float) -> IQMPS:
    """Bose-Hubbard ground state fixture (J=1) shipped with the package (tools/make_ground_states.py)."""
    path = os.path.join(DATA_DIR, f"bh_L{L}_d{d}_N{Npart}_U{U:g}.npz")
    if not os.path.exists(path):
        raise FileNotFoundError(f"no ground-state fixture {path}; generate it with tools/make_ground_states.py")
    return load_state(path)
