"""TEST INFRASTRUCTURE ONLY (see oracle/__init__.py): CPU restatement of the reference's site observables.

``expectation_values`` follows include/correlations.hpp:99-117 (`expectationValue`: psi.position(i), contract the site
tensor with the operator and its own conjugate; `expectationValues`: all sites).  Here the gauge is not moved: the left
and right environments are contracted explicitly, which gives the same number for any gauge.  Operators are the
boson-number-diagonal ones of include/BH_sites.h:129-171 (N, N(N-1), NN)."""
import numpy as np


def expectation_values(psi, op_diag):
    """[<psi| diag(op_diag) at site j |psi> for j = 0..L-1] (not divided by <psi|psi>, like the reference)."""
    A = psi.A
    L = len(A)
    op = np.asarray(op_diag, dtype=float)
    left = [np.ones((1, 1), dtype=complex)]
    for j in range(L):                       # left[j+1][a', a] = sum conj(A[l', s, a']) left[j][l', l] A[l, s, a]
        left.append(np.einsum("xy,xsa,ysb->ab", left[j], A[j].conj(), A[j]))
    right = [None] * (L + 1)
    right[L] = np.ones((1, 1), dtype=complex)
    for j in range(L - 1, -1, -1):
        right[j] = np.einsum("xsa,ysb,ab->xy", A[j].conj(), A[j], right[j + 1])
    out = np.zeros(L)
    for j in range(L):
        v = np.einsum("xy,xsa,s,ysb,ab->", left[j], A[j].conj(), op, A[j], right[j + 1])
        out[j] = v.real
    return out


def entanglement_entropy(psi):
    """von Neumann entropies of the L-1 bonds, include/correlations.hpp:119-148: position(i), SVD of the two-site
    wavefunction, S = -sum_{p > 1e-12} p ln p over the density-matrix eigenvalues p = sigma^2 (not renormalised).
    With the centre at site i the site tensor as a (left x physical) x right matrix has the same singular values as the
    two-site wavefunction (site i+1 is right-orthonormal), so a copy is gauged site by site and that matrix is decomposed."""
    phi = psi.copy()
    L = len(phi.A)
    out = np.zeros(L - 1)
    for i in range(1, L):
        phi.position(i)
        a = phi.A[i - 1]
        sv = np.linalg.svd(a.reshape(a.shape[0] * a.shape[1], a.shape[2]), compute_uv=False)
        p = sv ** 2
        p = p[p > 1e-12]
        out[i - 1] = float(-(p * np.log(p)).sum())
    return out


def correlation_function(psi, op1, i, op2, j):
    """<psi| Op1_i Op2_j |psi>, sites 1-based, operators as D x D matrices <t|O|s> (include/correlations.hpp:10-55).
    Follows the reference's contraction: for i == j the product Op1.Op2 on that site (:17-24, the reference returns its real
    part); for i != j the transfer-matrix chain between the two sites (:37-52) -- here with the full left and right
    environments instead of a gauge move, which gives the same number in any gauge."""
    A = psi.A
    L = len(A)
    D = A[0].shape[1]
    ops = {}
    if i == j:
        ops[i - 1] = np.asarray(op1, dtype=float) @ np.asarray(op2, dtype=float)
    else:
        ops[i - 1] = np.asarray(op1, dtype=float)
        ops[j - 1] = np.asarray(op2, dtype=float)
    E = np.ones((1, 1), dtype=complex)
    for k in range(L):
        O = ops.get(k, np.eye(D))
        E = np.einsum("xy,xta,ts,ysb->ab", E, A[k].conj(), O, A[k])
    return complex(E[0, 0])


def correlation_matrix(psi, op1, op2):
    """include/correlations.hpp:57-80."""
    L = len(psi.A)
    rho = np.zeros((L, L), dtype=complex)
    for i in range(1, L + 1):
        rho[i - 1, i - 1] = correlation_function(psi, op1, i, op2, i).real
        for j in range(i + 1, L + 1):
            c = correlation_function(psi, op1, i, op2, j)
            rho[i - 1, j - 1] = c
            rho[j - 1, i - 1] = np.conj(c)
    return rho


def correlation_term(psi, op1, op2):
    """Largest eigenvalue of the correlation matrix (include/correlations.hpp:82-97)."""
    return float(np.linalg.eigvalsh(correlation_matrix(psi, op1, op2))[-1])
