"""Oracle (test infrastructure): Bose-Hubbard ground states as charge-labelled MPS.

The reference gets psi_init / psi_target from ITensor DMRG (include/InitializeState.hpp:18-65:
H = -J sum (a_i adag_{i+1} + h.c.) + U/2 sum n(n-1), Npart bosons).  Here:
  * ``ground_state_ed``   exact diagonalisation in the Npart sector (small L), and
  * ``ground_state_dmrg`` a plain two-site DMRG on the charge-labelled dense MPS (large L).
"""
from __future__ import annotations

import itertools
import numpy as np
import scipy.sparse as sp
import scipy.sparse.linalg as spla

from .bh_mps import MPS, boson_ops, mps_from_statevector, product_state, TruncArgs, denmat_decomp


def _sector_basis(L, D, Npart):
    states = []

    def rec(prefix, left):
        if len(prefix) == L - 1:
            if 0 <= left < D:
                states.append(tuple(prefix) + (left,))
            return
        for n in range(min(D - 1, left) + 1):
            rec(prefix + [n], left - n)

    rec([], Npart)
    return states


def bh_hamiltonian_sector(L, D, Npart, J, U):
    states = _sector_basis(L, D, Npart)
    index = {s: i for i, s in enumerate(states)}
    rows, cols, vals = [], [], []
    for i, s in enumerate(states):
        diag = 0.5 * U * sum(n * (n - 1) for n in s)
        rows.append(i); cols.append(i); vals.append(diag)
        for b in range(L - 1):
            # -J adag_b a_{b+1}
            if s[b + 1] > 0 and s[b] < D - 1:
                t = list(s); t[b] += 1; t[b + 1] -= 1
                amp = -J * np.sqrt(s[b] + 1) * np.sqrt(s[b + 1])
                j = index[tuple(t)]
                rows.append(j); cols.append(i); vals.append(amp)
                rows.append(i); cols.append(j); vals.append(amp)
    H = sp.csr_matrix((vals, (rows, cols)), shape=(len(states), len(states)))
    return H, states


def ground_state_ed(L, D, Npart, J, U) -> MPS:
    H, states = bh_hamiltonian_sector(L, D, Npart, J, U)
    n = H.shape[0]
    if n <= 3000:
        w, v = np.linalg.eigh(H.toarray())
        vec = v[:, 0]
    else:
        w, v = spla.eigsh(H, k=1, which="SA", tol=1e-13)
        vec = v[:, 0]
    if vec[np.argmax(np.abs(vec))] < 0:
        vec = -vec
    full = np.zeros((D,) * L)
    for amp, s in zip(vec, states):
        full[s] = amp
    return mps_from_statevector(full, L, D)


# ----------------------------------------------------------------------------------------------
# two-site DMRG (dense tensors, charge labels kept by block decompositions)
# ----------------------------------------------------------------------------------------------
def _bh_mpo(L, D, J, U):
    op = boson_ops(D)
    I, A, Ad, K = op["Id"], op["A"], op["Adag"], 0.5 * U * op["N(N-1)"]
    # W[a, b, t, s]; state 0 = done, 3 = nothing yet, 1: A placed (needs Adag), 2: Adag placed (needs A)
    W = np.zeros((4, 4, D, D))
    W[0, 0] = I
    W[3, 3] = I
    W[3, 0] = K
    W[3, 1] = -J * A
    W[3, 2] = -J * Ad
    W[1, 0] = Ad
    W[2, 0] = A
    Ws = []
    for j in range(L):
        w = W
        if j == 0:
            w = w[3:4]
        if j == L - 1:
            w = w[:, 0:1]
        Ws.append(w)
    return Ws


def ground_state_dmrg(L, D, Npart, J, U, maxm_schedule=(10, 20, 50, 100, 200), cutoff=1e-9,
                      nsweeps=10, krylov=4, verbose=False) -> MPS:
    """Two-site DMRG with the reference's sweep schedule (InitializeState.hpp:52-57; the noise
    term is replaced by a larger Krylov space).  Real arithmetic; returns centre at site 1."""
    occ = [0] * L
    p = Npart
    for i in range(L - 1, -1, -1):         # :28-38 particles distributed from the right, one per site
        if p >= 1:
            occ[i] = 1
            p -= 1
    psi = product_state(occ, D)
    Ws = _bh_mpo(L, D, J, U)
    s = np.arange(D)
    # environments: Lenv[j] covers sites < j (shape [a', w, a]); Renv[j] covers sites > j
    Lenv = [None] * (L + 1)
    Renv = [None] * (L + 1)
    Lenv[0] = np.ones((1, 1, 1))
    Renv[L - 1] = np.ones((1, 1, 1))

    def grow_left(j):
        a = psi.A[j].real
        t = np.tensordot(Lenv[j], a, axes=(2, 0))                  # [a', w, s, r]
        t = np.tensordot(t, Ws[j], axes=([1, 2], [0, 3]))          # [a', r, w2, t]
        Lenv[j + 1] = np.tensordot(a, t, axes=([0, 1], [0, 3])).transpose(0, 2, 1)  # [r', w2, r]

    def grow_right(j):
        a = psi.A[j].real
        t = np.tensordot(a, Renv[j], axes=(2, 2))                  # [l, s, r', w]
        t = np.tensordot(t, Ws[j], axes=([3, 1], [1, 3]))          # [l, r', w1, t]
        Renv[j - 1] = np.tensordot(a, t, axes=([1, 2], [3, 1])).transpose(0, 2, 1)  # [l', w1, l]

    for j in range(L - 1, 1, -1):
        grow_right(j)
    energy = None
    for sw in range(nsweeps):
        maxm = maxm_schedule[min(sw, len(maxm_schedule) - 1)]
        args = TruncArgs(cutoff=cutoff, maxm=maxm)
        for direction in ("left", "right"):
            bonds = range(0, L - 1) if direction == "left" else range(L - 2, -1, -1)
            for b in bonds:
                A1, A2 = psi.A[b].real, psi.A[b + 1].real
                theta = np.tensordot(A1, A2, axes=(2, 0))          # [l,s1,s2,r]
                ql, qr = psi.q[b], psi.q[b + 2]
                mask = ((ql[:, None, None, None] + s[None, :, None, None] + s[None, None, :, None])
                        == qr[None, None, None, :])
                Le, Re, W1, W2 = Lenv[b], Renv[b + 1], Ws[b], Ws[b + 1]

                def heff(v):
                    x = v.reshape(theta.shape) * mask
                    t = np.tensordot(Le, x, axes=(2, 0))                        # [l', w, s1, s2, r]
                    t = np.tensordot(t, W1, axes=([1, 2], [0, 3]))              # [l', s2, r, w1, t1]
                    t = np.tensordot(t, W2, axes=([3, 1], [0, 3]))              # [l', r, t1, w2, t2]
                    t = np.tensordot(t, Re, axes=([1, 3], [2, 1]))              # [l', t1, t2, r']
                    return (t * mask).ravel()

                v0 = (theta * mask).ravel()
                nv = np.linalg.norm(v0)
                v0 = v0 / nv
                # small Lanczos
                Vs, alphas, betas = [v0], [], []
                w = heff(v0)
                for it in range(krylov):
                    a_ = float(Vs[-1] @ w)
                    alphas.append(a_)
                    w = w - a_ * Vs[-1] - (betas[-1] * Vs[-2] if betas else 0.0)
                    for vv in Vs:      # full reorthogonalisation
                        w = w - (vv @ w) * vv
                    b_ = np.linalg.norm(w)
                    if b_ < 1e-12 or it == krylov - 1:
                        break
                    betas.append(b_)
                    Vs.append(w / b_)
                    w = heff(Vs[-1])
                k = len(alphas)
                T = np.diag(alphas) + np.diag(betas[:k - 1], 1) + np.diag(betas[:k - 1], -1)
                ew, ev = np.linalg.eigh(T)
                energy = ew[0]
                vec = sum(ev[i, 0] * Vs[i] for i in range(k))
                vec = vec / np.linalg.norm(vec)
                th = (vec.reshape(theta.shape) * mask).astype(complex)
                A1n, A2n, qm = denmat_decomp(th, ql, qr, direction, args)
                psi.A[b], psi.A[b + 1], psi.q[b + 1] = A1n, A2n, qm
                if direction == "left":
                    if b < L - 2:
                        grow_left(b)
                else:
                    if b > 0:
                        grow_right(b + 1)
        if verbose:
            print(f"sweep {sw} maxm {maxm} E = {energy:.12f} dims {psi.bond_dims()}")
    # after the right->left half sweep the centre sits on site 1
    psi.llim, psi.rlim = 0, 2
    psi.normalize()
    psi.energy = energy
    return psi


def mps_energy(psi: MPS, J, U):
    """<psi|H|psi>/<psi|psi> via the MPO (test helper)."""
    Ws = _bh_mpo(psi.L, psi.D, J, U)
    E = np.ones((1, 1, 1), dtype=complex)
    N = np.ones((1, 1), dtype=complex)
    for j in range(psi.L):
        a = psi.A[j]
        t = np.tensordot(E, a, axes=(2, 0))
        t = np.tensordot(t, Ws[j], axes=([1, 2], [0, 3]))
        E = np.tensordot(a.conj(), t, axes=([0, 1], [0, 3])).transpose(0, 2, 1)
        n = np.tensordot(N, a, axes=(1, 0))
        N = np.tensordot(a.conj(), n, axes=([0, 1], [0, 1]))
    return (E[0, 0, 0] / N[0, 0]).real
