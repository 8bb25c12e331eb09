"""Self-consistency of the oracle beyond the reference's goldens: exact state-vector evolution, finite
differences at the reference's own tolerances (tests/GradientTests.cpp, tests/HessianTests.cpp), the
new_control cache contract (tests/SequencingTest.cpp) and the ITensor-semantics helpers."""
import numpy as np
import pytest

from oracle import bh_mps as ob, optimal_control as oo, ground_state as og


def dense_step(vec, L, D, J, ts, u_from, u_to, forward=True):
    """Exact state-vector version of BH_tDMRG::step (no truncation)."""
    sgn = 1.0 if forward else -1.0
    G = ob.bond_gate(D, J, sgn * ts).reshape(D, D, D, D)
    u1 = ob.u_phases(D, sgn * u_from, ts)
    u2 = ob.u_phases(D, sgn * u_to, ts)
    psi = vec.reshape((D,) * L).copy()
    for j in range(L):
        sh = [1] * L
        sh[j] = D
        psi = psi * u1.reshape(sh)
    for (i1, i2) in ob.gate_order(L):
        a, b = i1 - 1, i2 - 1
        psi = np.moveaxis(np.tensordot(G, psi, axes=([2, 3], [a, b])), [0, 1], [a, b])
    for j in range(L):
        sh = [1] * L
        sh[j] = D
        psi = psi * u2.reshape(sh)
    return psi / np.linalg.norm(psi)


@pytest.mark.parametrize("L,d,Np", [(4, 3, 4), (5, 4, 5)])
def test_step_matches_statevector(L, d, Np):
    D = d + 1
    psi = og.ground_state_ed(L, D, Np, 1.0, 2.5)
    st = ob.BHStepper(L, D, 1.0, 1e-2, ob.TruncArgs(cutoff=1e-15))
    vec = psi.to_dense()
    for (a, b, fwd) in [(3.0, 4.0, True), (4.0, 6.0, True), (6.0, 5.0, False)]:
        st.step(psi, a, b, fwd)
        vec = dense_step(vec, L, D, 1.0, 1e-2, a, b, fwd)
        ov = abs(np.vdot(vec, psi.to_dense()))
        assert abs(ov - 1.0) < 1e-12
        assert psi.check_charges() == 0.0
        assert psi.llim == 0 and psi.rlim == 2


def test_gate_is_unitary_and_number_conserving():
    D = 5
    G = ob.bond_gate(D, 1.0, 0.01)
    assert np.allclose(G @ G.conj().T, np.eye(D * D), atol=1e-14)
    n = np.arange(D)
    tot = (n[:, None] + n[None, :]).ravel()
    assert np.all(G[tot[:, None] != tot[None, :]] == 0)       # exact zeros, needed by the charge bookkeeping
    assert np.allclose(ob.bond_gate(D, 1.0, -0.01), G.conj().T, atol=1e-14)


def test_truncate_rule():
    P = np.array([0.5, 0.3, 0.1, 0.05, 0.03, 0.02, 1e-9, 1e-10, 1e-12])
    m, err, docut = ob.truncate(P, maxm=5000, minm=1, cutoff=1e-8)
    assert m == 6 and docut == pytest.approx((0.02 + 1e-9) / 2)
    m, err, docut = ob.truncate(P, maxm=4, minm=1, cutoff=1e-8)
    assert m == 4
    m, err, docut = ob.truncate(np.array([0.5, 0.25, 0.25]), maxm=2, minm=1, cutoff=0.0)
    assert m == 2 and docut == pytest.approx(0.25 + 1e-3 * 0.25)    # degenerate pair straddling the cut
    assert ob.truncate(np.array([0.3]))[0] == 1


def test_position_preserves_state_and_orthogonality():
    psi = og.ground_state_ed(6, 4, 6, 1.0, 3.0)
    ref = psi.to_dense()
    psi.position(4)
    assert psi.llim == 3 and psi.rlim == 5
    assert np.allclose(psi.to_dense(), ref, atol=1e-13)
    for j in range(3):                 # left-orthonormal
        a = psi.A[j].reshape(-1, psi.A[j].shape[2])
        assert np.allclose(a.conj().T @ a, np.eye(a.shape[1]), atol=1e-13)
    for j in range(4, 6):              # right-orthonormal
        a = psi.A[j].reshape(psi.A[j].shape[0], -1)
        assert np.allclose(a @ a.conj().T, np.eye(a.shape[0]), atol=1e-13)
    psi.position(1)
    assert np.allclose(psi.to_dense(), ref, atol=1e-13)


def test_overlaps_and_apply_K():
    L, D = 5, 5
    a = og.ground_state_ed(L, D, 5, 1.0, 2.0)
    b = og.ground_state_ed(L, D, 5, 1.0, 9.0)
    va, vb = a.to_dense().ravel(), b.to_dense().ravel()
    assert ob.overlap(a, b) == pytest.approx(np.vdot(va, vb), abs=1e-13)
    n = np.arange(D)
    occ = np.indices((D,) * L).reshape(L, -1)
    kdiag = (0.5 * occ * (occ - 1)).sum(axis=0)
    assert ob.overlap_K(a, b) == pytest.approx(np.vdot(va, kdiag * vb), abs=1e-12)
    kb = ob.apply_K(b, ob.TruncArgs(cutoff=1e-14))
    assert np.allclose(kb.to_dense().ravel(), kdiag * vb, atol=1e-7)
    assert kb.norm() == pytest.approx(np.linalg.norm(kdiag * vb), rel=1e-10)
    assert kb.check_charges() == 0.0


@pytest.fixture(scope="module")
def grad_problem():      # tests/GradientTests.cpp:23-46 with T=0.1
    L, Npart, d = 5, 5, 5
    D = d + 1
    J, cs, ce, T, ts = 1.0, 2.0, 12.0, 0.1, 1e-2
    N = int(T / ts + 1)
    psi_i = og.ground_state_ed(L, D, Npart, J, cs)
    psi_f = og.ground_state_ed(L, D, Npart, J, ce)
    st = ob.BHStepper(L, D, J, ts, ob.TruncArgs(cutoff=1e-8))
    basis = oo.build_chopped_sine_basis(oo.linspace(cs, ce, N), ts, T, 6)
    return N, psi_i, psi_f, st, basis


def test_gradient_vs_finite_differences(grad_problem):        # tests/GradientTests.cpp:106-146
    N, psi_i, psi_f, st, basis = grad_problem
    oc = oo.OptimalControl(psi_f, psi_i, st, N=N, gamma=0)
    u = list(np.random.default_rng(5).uniform(2, 10, N))
    g = np.array(oc.getAnalyticGradient(u))
    eps = 1e-5
    num = []
    for i in range(N):
        up, um = list(u), list(u)
        up[i] += eps
        um[i] -= eps
        num.append((oc.getCost(up) - oc.getCost(um)) / (2 * eps))
    num = np.array(num)
    rel = np.abs(g[1:-1] - num[1:-1]) / np.abs(num[1:-1])
    assert rel.max() < 1e-3                                    # the reference's tolerance, interior entries only
    assert g[0] / num[0] == pytest.approx(2.0, rel=1e-3)       # end points are twice the derivative
    # BFGS branch gives the same gradient (tests/SequencingTest.cpp:127-133)
    ob_ = oo.OptimalControl(psi_f, psi_i, st, N=N, gamma=0, BFGS=True)
    assert np.allclose(ob_.getAnalyticGradient(u), g, atol=1e-12)
    # GROUP gradient = Jacobian^T . GRAPE gradient
    og_ = oo.OptimalControl(psi_f, psi_i, st, basis=basis, gamma=0)
    c = list(np.random.default_rng(6).uniform(-1, 1, basis.getM()))
    gc = np.array(og_.getAnalyticGradient(c))
    ug = basis.convertControl(c)
    gu = np.array(oc.getAnalyticGradient(ug))
    assert np.allclose(gc, np.array(basis.getControlJacobian()).T @ gu, atol=1e-12)


def test_hessian_vs_finite_differences(grad_problem):          # tests/HessianTests.cpp:131-184
    N, psi_i, psi_f, st, basis = grad_problem
    oc = oo.OptimalControl(psi_f, psi_i, st, N=N, gamma=0)
    u = list(np.random.default_rng(7).uniform(2, 10, N))
    H = np.array(oc.getHessian(u))
    g0 = np.array(oc.getAnalyticGradient(u))
    eps = 1e-3
    Hn = np.zeros((N, N))
    for i in range(N):
        up = list(u)
        up[i] += eps
        Hn[:, i] = (np.array(oc.getAnalyticGradient(up)) - g0) / eps
    I = slice(1, N - 1)
    rel = np.abs(H[I, I] - Hn[I, I]) / np.abs(Hn[I, I])
    assert rel.max() < 5e-3
    assert np.allclose(H, H.T)
    assert np.all(H[0] == 0) and np.all(H[:, N - 1] == 0)     # end-point rows/columns are never filled


def test_new_control_cache_contract():                         # tests/SequencingTest.cpp
    L, D = 3, 4
    J, T, ts = 2.0, 0.2, 1e-2
    N = int(T / ts + 1)
    psi_i = og.ground_state_ed(L, D, 3, J, 2.0)
    psi_f = og.ground_state_ed(L, D, 3, J, 12.0)
    st = ob.BHStepper(L, D, J, ts, ob.TruncArgs(cutoff=1e-7))
    rng = np.random.default_rng(3)
    u = list(rng.uniform(5, 15, N))
    ref = oo.OptimalControl(psi_f, psi_i, st, N=N, gamma=0)
    c0, g0, H0 = ref.getCost(u), np.array(ref.getAnalyticGradient(u)), np.array(ref.getHessian(u))
    import itertools
    calls = {"c": lambda o, nc: o.getCost(u, nc), "g": lambda o, nc: np.array(o.getAnalyticGradient(u, nc)),
             "h": lambda o, nc: np.array(o.getHessian(u, nc))}
    for order in itertools.permutations("cgh"):                # all six call orders (:116-198)
        o = oo.OptimalControl(psi_f, psi_i, st, N=N, gamma=0)
        res = {}
        for k, name in enumerate(order):
            res[name] = calls[name](o, k == 0)
        assert abs(res["c"] - c0) < 1e-10
        assert np.max(np.abs(res["g"] - g0)) < 1e-10
        assert np.max(np.abs(res["h"] - H0)) < 1e-10
    # a different control with new_control=false returns the stale cost and gradient (:238-256) ...
    u2 = list(rng.uniform(2, 20, N))
    assert abs(ref.getCost(u2, False) - c0) < 1e-10
    assert np.max(np.abs(np.array(ref.getAnalyticGradient(u2, False)) - g0)) < 1e-10
    assert abs(ref.getCost(u2, True) - c0) > 1e-10
    # ... but a different Hessian, because rows are re-propagated with the passed control (:258-266)
    ref.getCost(u, True)
    ref.getAnalyticGradient(u, True)
    u3 = list(rng.uniform(1, 4, N))
    assert np.max(np.abs(np.array(ref.getHessian(u3, False)) - H0)) > 1e-10


def test_dmrg_matches_exact_diagonalisation():
    L, D, Np = 6, 5, 6
    for U in (2.5, 30.0):
        ed = og.ground_state_ed(L, D, Np, 1.0, U)
        dm = og.ground_state_dmrg(L, D, Np, 1.0, U, maxm_schedule=(10, 20, 50, 100), cutoff=1e-10, nsweeps=8)
        assert abs(ob.overlap(ed, dm)) == pytest.approx(1.0, abs=1e-7)
        assert dm.check_charges() == 0.0


def test_schedule_touches_every_bond_once():
    for L in (2, 3, 4, 5, 6, 7, 20):
        gates = ob.gate_order(L)
        assert sorted(gates) == [(i, i + 1) for i in range(1, L)]


def test_observables_restatement_on_exact_states():
    """oracle.observables against dense state vectors: <N_j> of a number-conserving state sums to the particle number,
    and every entry equals the expectation value computed from the full state vector."""
    import itertools
    from oracle import bh_mps as ob, observables as obs
    L, D, Np = 4, 3, 3
    rng = np.random.default_rng(5)
    confs = [c for c in itertools.product(range(D), repeat=L) if sum(c) == Np]
    vec = np.zeros((D,) * L, dtype=complex)
    for c in confs:
        vec[c] = rng.normal() + 1j * rng.normal()
    vec /= np.linalg.norm(vec)
    psi = ob.mps_from_statevector(vec.reshape(-1), L, D)
    n = np.arange(D, dtype=float)
    got = obs.expectation_values(psi, n)
    want = [float(sum(abs(vec[c]) ** 2 * c[j] for c in confs)) for j in range(L)]
    assert np.allclose(got, want, atol=1e-12)
    assert abs(got.sum() - Np) < 1e-12
    got2 = obs.expectation_values(psi, n * (n - 1))
    want2 = [float(sum(abs(vec[c]) ** 2 * c[j] * (c[j] - 1) for c in confs)) for j in range(L)]
    assert np.allclose(got2, want2, atol=1e-12)


def test_entropy_restatement_on_exact_states():
    """oracle.observables.entanglement_entropy against the Schmidt spectrum of the dense state vector."""
    import itertools
    from oracle import bh_mps as ob, observables as obs
    L, D, Np = 4, 3, 4
    rng = np.random.default_rng(11)
    vec = np.zeros((D,) * L, dtype=complex)
    for c in itertools.product(range(D), repeat=L):
        if sum(c) == Np:
            vec[c] = rng.normal() + 1j * rng.normal()
    vec /= np.linalg.norm(vec)
    psi = ob.mps_from_statevector(vec.reshape(-1), L, D)
    got = obs.entanglement_entropy(psi)
    for i in range(1, L):
        p = np.linalg.svd(vec.reshape(D ** i, D ** (L - i)), compute_uv=False) ** 2
        p = p[p > 1e-12]
        assert abs(got[i - 1] + (p * np.log(p)).sum()) < 1e-12


def test_correlation_function_restatement_against_dense_state_vector():
    """oracle.observables.correlation_function (include/correlations.hpp:10-55) against brute-force state-vector arithmetic."""
    from conftest import random_symmetric_mps
    from oracle import observables as oo
    from optimalcontrolmps_b200.api import site_operator
    L, D, Np = 4, 4, 4
    psi = random_symmetric_mps(L, D, Np, 8, 3)
    v = psi.to_dense()

    def dense(o1, i, o2, j):
        ops = [np.eye(D)] * L
        if i == j:
            ops[i - 1] = o1 @ o2
        else:
            ops[i - 1], ops[j - 1] = o1, o2
        w = v
        for k, O in enumerate(ops):
            w = np.moveaxis(np.tensordot(O, w, axes=(1, k)), 0, k)
        return np.vdot(v, w)

    for a, b in (("Adag", "A"), ("N", "N"), ("A", "Adag"), ("Id", "N(N-1)")):
        o1, o2 = site_operator(a, D), site_operator(b, D)
        for i in range(1, L + 1):
            for j in range(i, L + 1):
                assert abs(oo.correlation_function(psi, o1, i, o2, j) - dense(o1, i, o2, j)) < 1e-12
    rho = oo.correlation_matrix(psi, site_operator("Adag", D), site_operator("A", D))
    assert abs(np.trace(rho).real - Np) < 1e-12 and np.max(np.abs(rho - rho.conj().T)) < 1e-14


def test_oracle_reproduces_the_start_of_the_cfg5_quench_golden():
    """The committed golden of the cfg5 quench (tests/golden/make_golden_cfg5.py, 60 s of CPU) is checked here on its first 14 steps
    (bond dimensions up to ~45, well under a second each): guards the golden against drift of the oracle."""
    import os
    import bench
    z = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "golden_cfg5_quench.npz"))
    c = bench.CFG5
    L, D = c["L"], c["d"] + 1
    st = ob.BHStepper(L, D, c["J"], 5 * c["tstep"], ob.TruncArgs(cutoff=c["cutoff"], maxm=c["maxm"]))
    psi0 = ob.product_state([1] * L, D)
    psi = psi0.copy()
    for k in range(14):
        st.step(psi, 2.5, 2.5, True)
        assert list(psi.bond_dims()) == z["dims"][k + 1].tolist()
        assert abs(ob.overlap(psi0, psi) - complex(z["amp"][k + 1])) < 1e-12


def test_oracle_reproduces_the_start_of_the_cfg4_sweep_golden():
    """Same for the cfg4 sweep golden (tests/golden/make_golden_cfg4.py): the first 25 time points from the L=30 ground states."""
    import os
    import bench
    from optimalcontrolmps_b200.states import ground_state
    z = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "golden_cfg4_sweep.npz"))
    c = bench.CFG4
    D = c["d"] + 1
    conv = lambda h: ob.MPS(h.A, [np.asarray(x, dtype=np.int64) for x in h.q], 0, 2)
    psi = conv(ground_state(c["L"], c["d"], c["Npart"], c["U_i"]))
    target = conv(ground_state(c["L"], c["d"], c["Npart"], c["U_f"]))
    st = ob.BHStepper(c["L"], D, c["J"], c["tstep"], ob.TruncArgs(cutoff=c["cutoff"], maxm=c["maxm"]))
    u = z["u"]
    assert list(psi.bond_dims()) == z["psi_dims"][0].tolist()
    for k in range(24):
        st.step(psi, u[k], u[k + 1], True)
        assert list(psi.bond_dims()) == z["psi_dims"][k + 1].tolist()
        f = abs(ob.overlap(target, psi)) ** 2
        assert abs(f - z["fidelities"][k + 1]) <= 1e-12 * max(1.0, z["fidelities"][k + 1])
