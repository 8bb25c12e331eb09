"""Oracle (test infrastructure): Bose-Hubbard MPS time stepper, NumPy restatement.

Follows the reference ``src/BH_tDMRG.cpp`` (gate construction :18-58, U gates :74-108,
``step`` :111-125, ``doStep`` :127-230, ``propDeriv`` :10-14) and the ITensor v2 calls it
makes (``BondGate``, ``denmatDecomp``, ``MPS::position``/``orthMPS``, ``normalize``,
``overlap``/``overlapC``, ``exactApplyMPO``) as recalled in SURVEY.md appendix A.

Representation: ITensor's IQTensors are block sparse in the boson number.  Here every
site tensor is a dense ``complex128`` array ``A[l, s, r]`` and every bond carries an
integer charge label per index (= number of bosons to the left of the bond).  All
decompositions are done per charge block, exactly like ITensor's IQTensor code paths, so
entries that violate ``q_l + s == q_r`` are exact zeros throughout.

Sites are numbered 1..L in the public methods to make the code comparable with the
reference line by line.
"""
from __future__ import annotations

import numpy as np

MIN_CUT = 1e-16   # ITensor default cutoff when "Cutoff" is not given
MAX_M = 5000      # ITensor default "Maxm"


# ----------------------------------------------------------------------------------------------
# operators (include/BH_sites.h:129-171)
# ----------------------------------------------------------------------------------------------
def boson_ops(D: int):
    """Matrix elements <t|Op|s> for a site with states |0>..|D-1>  (BH_sites.h:129-171)."""
    n = np.arange(D, dtype=float)
    A = np.zeros((D, D))
    for j in range(1, D):
        A[j - 1, j] = np.sqrt(j)          # <j-1|A|j> = sqrt(j)      :136-141
    return {
        "A": A,
        "Adag": A.T.copy(),               # <j|Adag|j-1> = sqrt(j)   :143-148
        "N": np.diag(n),                  # :129-134
        "N(N-1)": np.diag(n * n - n),     # :150-155
        "NN": np.diag(n * n),             # :157-162
        "Id": np.eye(D),
    }


def bond_hamiltonian(D: int, J: float) -> np.ndarray:
    """h[(t1,t2),(s1,s2)] = -J (A_1 Adag_2 + Adag_1 A_2)   (src/BH_tDMRG.cpp:31-32)."""
    op = boson_ops(D)
    h = -J * (np.kron(op["A"], op["Adag"]) + np.kron(op["Adag"], op["A"]))
    return h


def bond_gate(D: int, J: float, tau: float, order: int = 100) -> np.ndarray:
    """exp(-i tau h) by ITensor BondGate's Horner recursion (BH_tDMRG.cpp:35-36)."""
    h = bond_hamiltonian(D, J).astype(complex)
    unit = np.eye(D * D, dtype=complex)
    x = h * (-1j * tau)
    term = x.copy()
    gate = unit
    for ordr in range(order, 0, -1):
        term = term / ordr
        gate = unit + term
        term = gate @ x
    return gate


def u_phases(D: int, U: float, tstep: float) -> np.ndarray:
    """exp(-i/4 U tstep n(n-1)), n=0..D-1   (src/BH_tDMRG.cpp:84-88)."""
    n = np.arange(D, dtype=float)
    return np.exp(-0.25j * U * tstep * n * (n - 1.0))


def k_diag(D: int) -> np.ndarray:
    """Diagonal of the propagator derivative 1/2 N(N-1)  (src/BH_tDMRG.cpp:10-14)."""
    n = np.arange(D, dtype=float)
    return 0.5 * n * (n - 1.0)


# ----------------------------------------------------------------------------------------------
# ITensor truncate()
# ----------------------------------------------------------------------------------------------
def truncate(P, maxm=MAX_M, minm=1, cutoff=MIN_CUT, absolute_cutoff=False, do_rel_cutoff=False):
    """P: eigenvalues sorted descending.  Returns (m, truncerr, docut)."""
    P = np.array(P, dtype=float)
    origm = len(P)
    n = origm - 1
    docut = 0.0
    # zero out trailing negative weight
    for zn in range(n, -1, -1):
        if P[zn] >= 0:
            break
        P[zn] = 0.0
    if origm == 1:
        return 1, 0.0, P[0] / 2.0
    truncerr = 0.0
    while n >= maxm:
        truncerr += P[n]
        n -= 1
    if absolute_cutoff:
        while n >= minm and P[n] < cutoff:
            truncerr += P[n]
            n -= 1
    else:
        scale = 1.0
        if do_rel_cutoff:
            scale = float(P.sum())
            if scale == 0.0:
                scale = 1.0
        while n >= minm and truncerr + P[n] < cutoff * scale:
            truncerr += P[n]
            n -= 1
        truncerr = 0.0 if P[0] == 0 else truncerr / scale
    if n < 0:
        n = 0
    m = n + 1
    if m < origm:
        docut = (P[m] + P[m - 1]) / 2.0
        if abs(P[m] - P[m - 1]) < 1e-3 * P[m - 1]:
            docut += 1e-3 * P[m - 1]
    return m, truncerr, docut


class TruncArgs:
    """The ITensor ``Args{"Cutoff=",x,"Maxm=",y}`` the stepper is built with.

    ``cutoff``/``maxm`` are ``None`` when the key was not passed (the reference's tests pass
    only Cutoff, tests/CostTests.cpp:41)."""

    def __init__(self, cutoff=None, maxm=None, rel_cutoff=False):
        self.cutoff = cutoff
        self.maxm = maxm
        self.rel_cutoff = rel_cutoff

    @property
    def cutoff_or_default(self):
        return MIN_CUT if self.cutoff is None else self.cutoff

    @property
    def maxm_or_default(self):
        return MAX_M if self.maxm is None else self.maxm


# ----------------------------------------------------------------------------------------------
# MPS container
# ----------------------------------------------------------------------------------------------
class MPS:
    """Dense-with-charge-labels MPS.  ``A[j]`` (0-based) has shape (chi_j, D, chi_{j+1});
    ``q[b]`` are the charges of bond b (b=0..L).  ``llim``/``rlim`` are ITensor's
    ``l_orth_lim_``/``r_orth_lim_`` (1-based sites)."""

    def __init__(self, A, q, llim=0, rlim=2):
        self.A = [np.ascontiguousarray(a, dtype=complex) for a in A]
        self.q = [np.asarray(x, dtype=np.int64).copy() for x in q]
        self.L = len(self.A)
        self.D = self.A[0].shape[1]
        self.llim = llim
        self.rlim = rlim
        assert len(self.q) == self.L + 1
        for j, a in enumerate(self.A):
            assert a.shape == (len(self.q[j]), self.D, len(self.q[j + 1])), (j, a.shape)

    def copy(self):
        return MPS([a.copy() for a in self.A], [x.copy() for x in self.q], self.llim, self.rlim)

    def bond_dims(self):
        return [len(x) for x in self.q]

    def touch(self, i):
        """ITensor ``Aref(i)``: widen the orthogonality limits."""
        if self.llim > i - 1:
            self.llim = i - 1
        if self.rlim < i + 1:
            self.rlim = i + 1

    def check_charges(self, tol=0.0):
        """max |entry| that violates q_l + s == q_r (must be exactly 0)."""
        worst = 0.0
        s = np.arange(self.D)
        for j, a in enumerate(self.A):
            ok = (self.q[j][:, None, None] + s[None, :, None]) == self.q[j + 1][None, None, :]
            bad = np.abs(a[~ok])
            if bad.size:
                worst = max(worst, float(bad.max()))
        return worst

    def to_dense(self):
        """Full state vector (small L only)."""
        v = self.A[0]
        for a in self.A[1:]:
            v = np.tensordot(v, a, axes=(v.ndim - 1, 0))
        return v.reshape(v.shape[1:-1])

    # ---- norm / normalize ----
    def ortho_center(self):
        assert self.llim + 2 == self.rlim, "MPS has no single orthogonality centre"
        return self.llim + 1

    def norm(self):
        return float(np.linalg.norm(self.A[self.ortho_center() - 1]))

    def normalize(self):
        c = self.ortho_center() - 1
        nrm = float(np.linalg.norm(self.A[c]))
        self.A[c] = self.A[c] / nrm
        return nrm

    # ---- centre moves: ITensor position() via orthMPS ----
    def position(self, i, cutoff=MIN_CUT, maxm=MAX_M):
        while self.llim < i - 1:
            if self.llim < 0:
                self.llim = 0
            b = self.llim + 1
            self._orth(b, "left", cutoff, maxm)
            self.llim += 1
            if self.rlim < self.llim + 2:
                self.rlim = self.llim + 2
        while self.rlim > i + 1:
            if self.rlim > self.L + 1:
                self.rlim = self.L + 1
            b = self.rlim - 2
            self._orth(b, "right", cutoff, maxm)
            self.rlim -= 1
            if self.llim > self.rlim - 2:
                self.llim = self.rlim - 2

    def _orth(self, b, direction, cutoff, maxm):
        """orthMPS on sites (b, b+1): block SVD of one site tensor, S.V pushed to the other."""
        D = self.D
        s = np.arange(D)
        if direction == "left":
            a = self.A[b - 1]
            chil, _, chir = a.shape
            X = a.reshape(chil * D, chir)
            rowq = (self.q[b - 1][:, None] + s[None, :]).ravel()
            colq = self.q[b]
            U, C, newq = _block_svd(X, rowq, colq, side="cols", cutoff=cutoff, maxm=maxm)
            # X ~= U @ C ; U (n x k) orthonormal columns, C (k x chir)
            self.A[b - 1] = U.reshape(chil, D, len(newq))
            nxt = self.A[b]
            self.A[b] = np.tensordot(C, nxt, axes=(1, 0))
            self.q[b] = newq
        else:
            a = self.A[b]                       # site b+1
            chil, _, chir = a.shape
            X = a.reshape(chil, D * chir)
            rowq = self.q[b]
            colq = (self.q[b + 1][None, :] - s[:, None]).ravel()
            # need right-orthonormal rows: decompose X^T
            U, C, newq = _block_svd(X.T, colq, rowq, side="cols", cutoff=cutoff, maxm=maxm)
            # X^T ~= U @ C  ->  X ~= C^T @ U^T
            self.A[b] = U.T.reshape(len(newq), D, chir)
            prv = self.A[b - 1]
            self.A[b - 1] = np.tensordot(prv, C.T, axes=(2, 0))
            self.q[b] = newq


def _block_svd(X, rowq, colq, side, cutoff, maxm):
    """Block SVD X = U S Vh per charge (rows and columns with equal charge label form a
    block), global sort + ITensor truncate, per-block filter ``sigma^2 > docut``.
    Returns U (n x k, orthonormal columns), C = S.Vh (k x m), and the new charges."""
    n, m = X.shape
    blocks = []
    alleig = []
    for q in np.unique(colq):
        cidx = np.nonzero(colq == q)[0]
        ridx = np.nonzero(rowq == q)[0]
        if ridx.size == 0:
            continue
        Xq = X[np.ix_(ridx, cidx)]
        u, sv, vh = np.linalg.svd(Xq, full_matrices=False)
        blocks.append((q, ridx, cidx, u, sv, vh))
        alleig.extend((sv * sv).tolist())
    alleig = np.sort(np.array(alleig))[::-1]
    mkeep, _, docut = truncate(alleig, maxm=maxm, minm=1, cutoff=cutoff)
    Ucols, Crows, newq = [], [], []
    total = 0
    for (q, ridx, cidx, u, sv, vh) in blocks:
        keep = int(np.sum(sv * sv > docut))
        if keep == 0:
            continue
        total += keep
        Uq = np.zeros((n, keep), dtype=complex)
        Uq[ridx, :] = u[:, :keep]
        Cq = np.zeros((keep, m), dtype=complex)
        Cq[:, cidx] = sv[:keep, None] * vh[:keep, :]
        Ucols.append(Uq)
        Crows.append(Cq)
        newq.extend([q] * keep)
    if total == 0:  # zero tensor: keep one arbitrary state (ITensor does the same)
        q, ridx, cidx, u, sv, vh = blocks[0]
        Uq = np.zeros((n, 1), dtype=complex)
        Uq[ridx, 0] = u[:, 0]
        Cq = np.zeros((1, m), dtype=complex)
        Ucols, Crows, newq = [Uq], [Cq], [q]
    return np.concatenate(Ucols, axis=1), np.concatenate(Crows, axis=0), np.array(newq, dtype=np.int64)


def _block_eig_basis(blocks_rho, cutoff, maxm, rel_cutoff=False):
    """Common tail of denmatDecomp/diagHermitian: ``blocks_rho`` is a list of
    (q, idx, rho_q).  Returns list of (q, idx, V_q_kept) with global truncation."""
    evs = []
    alleig = []
    for (q, idx, rho) in blocks_rho:
        w, v = np.linalg.eigh(rho)
        w = w[::-1]
        v = v[:, ::-1]
        evs.append((q, idx, w, v))
        alleig.extend(w.tolist())
    alleig = np.sort(np.array(alleig))[::-1]
    m, truncerr, docut = truncate(alleig, maxm=maxm, minm=1, cutoff=cutoff, do_rel_cutoff=rel_cutoff)
    out = []
    total = 0
    for (q, idx, w, v) in evs:
        keep = int(np.sum(w > docut))
        if keep == 0:
            continue
        total += keep
        out.append((q, idx, v[:, :keep]))
    if total == 0:
        q, idx, w, v = evs[0]
        out = [(q, idx, v[:, :1])]
    return out


def denmat_decomp(theta, ql, qr, direction, args: TruncArgs):
    """ITensor ``denmatDecomp`` on the two-site tensor theta[l,t1,t2,r].
    Returns (A1, A2, qmid)."""
    chil, D, _, chir = theta.shape
    s = np.arange(D)
    X = theta.reshape(chil * D, D * chir)
    rowq = (ql[:, None] + s[None, :]).ravel()                 # charge of the middle bond seen from the left
    colq = (qr[None, :] - s[:, None]).ravel()                 # ... seen from the right
    cutoff = args.cutoff_or_default
    maxm = args.maxm_or_default
    blocks = []
    if direction == "left":
        for q in np.unique(rowq):
            ridx = np.nonzero(rowq == q)[0]
            cidx = np.nonzero(colq == q)[0]
            if cidx.size == 0:
                Xq = np.zeros((ridx.size, 1), dtype=complex)
            else:
                Xq = X[np.ix_(ridx, cidx)]
            blocks.append((q, ridx, Xq @ Xq.conj().T))
        kept = _block_eig_basis(blocks, cutoff, maxm, args.rel_cutoff)
        k = sum(v.shape[1] for (_, _, v) in kept)
        U = np.zeros((chil * D, k), dtype=complex)
        qmid = np.zeros(k, dtype=np.int64)
        c = 0
        for (q, idx, v) in kept:
            U[idx, c:c + v.shape[1]] = v
            qmid[c:c + v.shape[1]] = q
            c += v.shape[1]
        A1 = U.reshape(chil, D, k)
        A2 = (U.conj().T @ X).reshape(k, D, chir)
    else:
        for q in np.unique(colq):
            cidx = np.nonzero(colq == q)[0]
            ridx = np.nonzero(rowq == q)[0]
            if ridx.size == 0:
                Xq = np.zeros((1, cidx.size), dtype=complex)
            else:
                Xq = X[np.ix_(ridx, cidx)]
            blocks.append((q, cidx, Xq.conj().T @ Xq))
        kept = _block_eig_basis(blocks, cutoff, maxm, args.rel_cutoff)
        k = sum(v.shape[1] for (_, _, v) in kept)
        V = np.zeros((D * chir, k), dtype=complex)
        qmid = np.zeros(k, dtype=np.int64)
        c = 0
        for (q, idx, v) in kept:
            V[idx, c:c + v.shape[1]] = v
            qmid[c:c + v.shape[1]] = q
            c += v.shape[1]
        A2 = V.conj().T.reshape(k, D, chir)
        A1 = (X @ V).reshape(chil, D, k)
    return A1, A2, qmid


# ----------------------------------------------------------------------------------------------
# the stepper (src/BH_tDMRG.cpp)
# ----------------------------------------------------------------------------------------------
def gate_order(L: int):
    """Bond list (i1, i2) in the order of initJGates (src/BH_tDMRG.cpp:28-57)."""
    gates = [(i, i + 1) for i in range(1, L, 2)]
    offset = 2 if L % 2 == 0 else 1
    gates += [(i, i + 1) for i in range(L - offset, 0, -2)]
    return gates


class BHStepper:
    """Restatement of class BH_tDMRG (include/BH_tDMRG.hpp:16-40)."""

    def __init__(self, L, D, J, tstep, args: TruncArgs):
        self.L, self.D, self.J = L, D, J
        self.args = args
        self.set_tstep(tstep)

    def set_tstep(self, tstep):                      # :61-65
        self.tstep = tstep
        self.G_fwd = bond_gate(self.D, self.J, +tstep)
        self.G_bwd = bond_gate(self.D, self.J, -tstep)
        self.gates = gate_order(self.L)

    def get_tstep(self):
        return self.tstep

    def step(self, psi: MPS, u_from, u_to, forward=True):   # :111-125
        if forward:
            u1 = u_phases(self.D, u_from, self.tstep)
            u2 = u_phases(self.D, u_to, self.tstep)
            self._do_step(psi, u1, u2, self.G_fwd)
        else:
            u1 = u_phases(self.D, -u_from, self.tstep)
            u2 = u_phases(self.D, -u_to, self.tstep)
            self._do_step(psi, u1, u2, self.G_bwd)

    def _do_step(self, psi: MPS, u1, u2, G):               # :127-230
        L, D = self.L, self.D
        args = self.args
        G4 = G.reshape(D, D, D, D)                     # [t1,t2,s1,s2]
        if L % 2 != 0:                                 # :133-136 lonely U gate on the last site
            psi.touch(L)
            psi.A[L - 1] = psi.A[L - 1] * u1[None, :, None]
        moving_from_left = True
        gates = self.gates
        for gi, (i1, i2) in enumerate(gates):
            psi.touch(i1)
            psi.touch(i2)
            theta = np.tensordot(psi.A[i1 - 1], psi.A[i2 - 1], axes=(2, 0))   # [l,s1,s2,r]
            if moving_from_left:                       # :150-155
                theta = theta * (u1[None, :, None, None] * u1[None, None, :, None])
                theta = np.einsum("tuab,labr->ltur", G4, theta, optimize=True)
                if i2 == L and L % 2 == 0:
                    theta = theta * u2[None, None, :, None]
            else:                                      # :159
                theta = np.einsum("tuab,labr->ltur", G4, theta, optimize=True)
                theta = theta * (u2[None, :, None, None] * u2[None, None, :, None])
            ql, qr = psi.q[i1 - 1], psi.q[i2]
            if gi + 1 < len(gates):
                ni1, ni2 = gates[gi + 1]
                if ni1 >= i2:                          # :173-188
                    A1, A2, qm = denmat_decomp(theta, ql, qr, "left", args)
                    psi.A[i1 - 1], psi.A[i2 - 1], psi.q[i1] = A1, A2, qm
                    psi.llim = i1
                    if psi.rlim < i1 + 2:
                        psi.rlim = i1 + 2
                    psi.touch(i1 + 1)
                    nrm = np.linalg.norm(psi.A[i1])
                    if nrm > 1e-16:
                        psi.A[i1] = psi.A[i1] / nrm
                    psi.position(ni1)
                if ni1 < i2:                           # :189-199
                    A1, A2, qm = denmat_decomp(theta, ql, qr, "right", args)
                    psi.A[i1 - 1], psi.A[i2 - 1], psi.q[i1] = A1, A2, qm
                    if psi.llim > i1 - 1:
                        psi.llim = i1 - 1
                    psi.rlim = i1 + 1
                    psi.touch(i1)
                    nrm = np.linalg.norm(psi.A[i1 - 1])
                    if nrm > 1e-16:
                        psi.A[i1 - 1] = psi.A[i1 - 1] / nrm
                    psi.position(ni2)
                if i2 == ni1 or i1 == ni2:             # :200-204
                    moving_from_left = False
            else:                                      # :206-218
                A1, A2, qm = denmat_decomp(theta, ql, qr, "right", args)
                psi.A[i1 - 1], psi.A[i2 - 1], psi.q[i1] = A1, A2, qm
                psi.llim = i1 - 1
                psi.rlim = i1 + 1
                psi.touch(i1)
                nrm = np.linalg.norm(psi.A[i1 - 1])
                if nrm > 1e-16:
                    psi.A[i1 - 1] = psi.A[i1 - 1] / nrm
                psi.position(1)
        psi.touch(1)                                    # :222-223 lonely U gate on site 1
        psi.A[0] = psi.A[0] * u2[None, :, None]
        psi.normalize()                                 # :228


# ----------------------------------------------------------------------------------------------
# overlaps and exactApplyMPO
# ----------------------------------------------------------------------------------------------
def overlap(a: MPS, b: MPS) -> complex:
    """<a|b>, first argument conjugated (ITensor overlapC(a,b))."""
    E = np.ones((1, 1), dtype=complex)
    for Aa, Ab in zip(a.A, b.A):
        T = np.tensordot(E, Ab, axes=(1, 0))                   # [ra_prev, s, rb]
        E = np.tensordot(Aa.conj(), T, axes=([0, 1], [0, 1]))  # [ra, rb]
    return complex(E[0, 0])


def overlap_K(a: MPS, b: MPS) -> complex:
    """<a|K|b> with K = sum_j 1/2 n_j(n_j-1)  (overlapC(a, propDeriv, b))."""
    kd = k_diag(a.D)
    E0 = np.ones((1, 1), dtype=complex)     # K not applied yet
    E1 = np.zeros((1, 1), dtype=complex)    # K applied on an earlier site
    for Aa, Ab in zip(a.A, b.A):
        T0 = np.tensordot(E0, Ab, axes=(1, 0))
        T1 = np.tensordot(E1, Ab, axes=(1, 0)) + T0 * kd[None, :, None]
        E0 = np.tensordot(Aa.conj(), T0, axes=([0, 1], [0, 1]))
        E1 = np.tensordot(Aa.conj(), T1, axes=([0, 1], [0, 1]))
    return complex(E1[0, 0])


def _k_mpo(L, D):
    """W_j[a, b, s] for K = sum_j k_j, bond dimension 2 (state 1 = not applied yet)."""
    kd = k_diag(D)
    Ws = []
    for j in range(L):
        W = np.zeros((2, 2, D))
        W[0, 0, :] = 1.0
        W[1, 1, :] = 1.0
        W[1, 0, :] = kd
        if j == 0:
            W = W[1:2, :, :]
        if j == L - 1:
            W = W[:, 0:1, :]
        Ws.append(W)
    return Ws


def apply_K(psi: MPS, args: TruncArgs) -> MPS:
    """ITensor v2.1 ``exactApplyMPO(K, psi, args)`` (density-matrix algorithm).
    Result is unnormalised with the orthogonality centre at site 1."""
    L, D = psi.L, psi.D
    if L == 1:
        out = psi.copy()
        out.A[0] = out.A[0] * k_diag(D)[None, :, None]
        return out
    Ws = _k_mpo(L, D)
    cutoff = 1e-13 if args.cutoff is None else args.cutoff
    maxm_set = args.maxm is not None
    s = np.arange(D)
    # B_j[(l,a), s, (r,b)] : the exact product K|psi>
    B = []
    for j in range(L):
        a = psi.A[j]
        W = Ws[j]
        t = a[:, None, :, :, None] * W.transpose(0, 2, 1)[None, :, :, None, :]   # [l,a,s,r,b]
        B.append(t)
    # left environments G_j[(r',b'),(r,b)] = <x'|x> of the left blocks
    G = [None] * L
    g = np.ones((1, 1, 1, 1), dtype=complex)                     # [l',a',l,a]
    for j in range(L - 1):
        t = B[j]
        # g[l',a',l,a] conj(t[l',a',s,r',b']) t[l,a,s,r,b]
        tmp = np.tensordot(g, t, axes=([2, 3], [0, 1]))          # [l',a',s,r,b]
        g = np.tensordot(t.conj(), tmp, axes=([0, 1, 2], [0, 1, 2]))  # [r',b',r,b]
        G[j + 1] = g
    newA = [None] * L
    newq = [None] * (L + 1)
    newq[L] = psi.q[L].copy()
    newq[0] = psi.q[0].copy()
    # O[l, a, s, m]
    O = B[L - 1].reshape(B[L - 1].shape[0], B[L - 1].shape[1], D, 1)
    qm = psi.q[L]                                                 # charges of index m
    for j in range(L - 1, 0, -1):                                 # 0-based site j = L-1 .. 1
        chil, na = O.shape[0], O.shape[1]
        m = O.shape[3]
        Om = O.reshape(chil * na, D * m)
        g = G[j]                                                  # [l',a',l,a]
        Gm = g.reshape(chil * na, chil * na)                      # Gm[x', x]
        colq = (qm[None, :] - s[:, None]).ravel()
        if maxm_set:
            maxm = args.maxm
        else:
            maxm = psi.A[j].shape[0] * (Ws[j].shape[0])
        blocks = []
        for q in np.unique(colq):
            cidx = np.nonzero(colq == q)[0]
            Oq = Om[:, cidx]
            rho = Oq.T @ Gm.T @ Oq.conj()
            rho = 0.5 * (rho + rho.conj().T)
            blocks.append((q, cidx, rho))
        kept = _block_eig_basis(blocks, cutoff, maxm, args.rel_cutoff)
        k = sum(v.shape[1] for (_, _, v) in kept)
        V = np.zeros((D * m, k), dtype=complex)
        qk = np.zeros(k, dtype=np.int64)
        c = 0
        for (q, idx, v) in kept:
            V[idx, c:c + v.shape[1]] = v
            qk[c:c + v.shape[1]] = q
            c += v.shape[1]
        newA[j] = V.T.reshape(k, D, m)
        newq[j] = qk
        C = (Om @ V.conj()).reshape(chil, na, k)                  # carry [l,a,k]
        # O_new[l',a',s',k] = sum_{l,a} B_{j-1}[l',a',s',l,a] C[l,a,k]
        O = np.tensordot(B[j - 1], C, axes=([3, 4], [0, 1]))
        qm = qk
    newA[0] = O.reshape(1, D, O.shape[3]) if O.shape[0] * O.shape[1] == 1 else None
    assert newA[0] is not None
    res = MPS(newA, newq, llim=0, rlim=2)
    return res


# ----------------------------------------------------------------------------------------------
# construction helpers
# ----------------------------------------------------------------------------------------------
def mps_from_statevector(psi, L, D, tol=1e-14) -> MPS:
    """Exact (up to ``tol``) MPS with charge labels from a number-conserving state
    ``psi[s_1,...,s_L]``; orthogonality centre ends at site 1."""
    s = np.arange(D)
    psi = np.asarray(psi, dtype=complex)
    big = np.unravel_index(int(np.argmax(np.abs(psi))), (D,) * L)
    tot = int(sum(big))                                    # particle number of the sector
    gnorm = float(np.linalg.norm(psi))
    M = psi.reshape(1, -1)
    ql = np.array([0], dtype=np.int64)
    A, q = [], [ql]
    for j in range(L - 1):
        chil = M.shape[0]
        rest = M.shape[1] // D
        X = M.reshape(chil * D, rest)
        rowq = (ql[:, None] + s[None, :]).ravel()
        Ucols, Crows, newq = [], [], []
        for qq in np.unique(rowq):
            ridx = np.nonzero(rowq == qq)[0]
            Xq = X[ridx]
            if not np.any(Xq):
                continue
            u, sv, vh = np.linalg.svd(Xq, full_matrices=False)
            keep = int(np.sum(sv > tol * gnorm))
            if keep == 0:
                continue
            Uq = np.zeros((chil * D, keep), dtype=complex)
            Uq[ridx] = u[:, :keep]
            Ucols.append(Uq)
            Crows.append(sv[:keep, None] * vh[:keep])
            newq.extend([qq] * keep)
        U = np.concatenate(Ucols, axis=1)
        M = np.concatenate(Crows, axis=0)
        newq = np.array(newq, dtype=np.int64)
        A.append(U.reshape(chil, D, len(newq)))
        q.append(newq)
        ql = newq
    chil = M.shape[0]
    A.append(M.reshape(chil, D, 1))
    q.append(np.array([tot], dtype=np.int64))
    mps = MPS(A, q, llim=L - 1, rlim=L + 1)
    mps.position(1)
    return mps


def product_state(occ, D) -> MPS:
    """|n_1 n_2 ... n_L> as an MPS."""
    L = len(occ)
    A, q = [], [np.array([0], dtype=np.int64)]
    tot = 0
    for n in occ:
        a = np.zeros((1, D, 1), dtype=complex)
        a[0, n, 0] = 1.0
        A.append(a)
        tot += n
        q.append(np.array([tot], dtype=np.int64))
    return MPS(A, q, llim=0, rlim=2)
