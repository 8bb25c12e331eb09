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
