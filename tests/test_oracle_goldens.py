"""Pins the oracle to every golden vector the reference's own tests hold for the path
(SURVEY.md section 8c): tests/CostTests.cpp, tests/ControlBasisTests.cpp, tests/SiteSetTests.cpp."""
import numpy as np
import pytest

from oracle import bh_mps as ob, optimal_control as oo, ground_state as og

# tests/CostTests.cpp:23-46 -- L=5, Npart=5, d=5, J=1, U 2 -> 50, T=0.1, tstep=0.01, M=5, Cutoff 1e-8
FID_LIN = [0.214338, 0.214325, 0.215126, 0.217281, 0.221019, 0.22621, 0.232328, 0.238484, 0.243617, 0.246862, 0.24801]       # :75
FID_ONE = [0.214338, 0.214233, 0.213919, 0.213398, 0.212672, 0.211744, 0.210618, 0.2093, 0.207796, 0.206112, 0.204256]      # :90
FID_GRP = [0.214338, 0.21411, 0.216706, 0.222581, 0.229759, 0.23623, 0.242512, 0.249913, 0.256515, 0.259334, 0.259687]      # :124
# The reference's goldens come from DMRG ground states (10 sweeps, niter 2, cutoff 1e-9) whose inexactness is
# visible already at t=0 (5.8e-6); with exact ground states the agreement is 1e-5 (SURVEY.md 8c "Verdict").
TOL = 1e-5


@pytest.fixture(scope="module")
def cost_problem():
    L, Npart, d = 5, 5, 5
    D = d + 1
    J, cs, ce, T, ts = 1.0, 2.0, 50.0, 0.1, 1e-2
    N = int(T / ts + 1)
    M = 5
    psi_i = og.ground_state_ed(L, D, Npart, J, cs)
    psi_f = og.ground_state_ed(L, D, Npart, J, ce)
    st = ob.BHStepper(L, D, J, ts, ob.TruncArgs(cutoff=1e-8))
    u0 = oo.linspace(cs, ce, N)
    basis = oo.build_chopped_sine_basis(u0, ts, T, M)
    grape = oo.OptimalControl(psi_f, psi_i, st, N=N, gamma=0)
    group = oo.OptimalControl(psi_f, psi_i, st, basis=basis, gamma=0)
    return N, M, grape, group


def test_grape_fidelities(cost_problem):                      # tests/CostTests.cpp:67-99
    N, M, grape, _ = cost_problem
    grape.setGamma(0)
    u = oo.linspace(2.0, 50.0, N)
    assert len(u) == N == 11
    cost = grape.getCost(u)
    fid = grape.getFidelityForAllT(u, False)
    assert abs(cost - 0.375995) < TOL
    assert np.max(np.abs(np.array(fid) - FID_LIN)) < TOL
    cost2 = grape.getCost([1.0] * N)
    fid2 = grape.getFidelityForAllT([1.0] * N, False)
    assert abs(cost2 - 0.397872) < TOL
    assert np.max(np.abs(np.array(fid2) - FID_ONE)) < TOL


def test_group_fidelities(cost_problem):                      # :102-133
    N, M, _, group = cost_problem
    group.setGamma(0)
    cost = group.getCost([0.0] * M)
    fid = group.getFidelityForAllT([0.0] * M, False)
    assert abs(cost - 0.375995) < TOL
    assert np.max(np.abs(np.array(fid) - FID_LIN)) < TOL
    c2 = oo.linspace(0, 7, M)
    cost2 = group.getCost(c2)
    fid2 = group.getFidelityForAllT(c2, False)
    assert abs(cost2 - 0.370157) < TOL
    assert np.max(np.abs(np.array(fid2) - FID_GRP)) < TOL


def test_regularization(cost_problem):                        # :136-203
    N, M, grape, group = cost_problem
    grape.setGamma(1)
    assert abs(grape.getCost(oo.linspace(2.0, 50.0, N)) - 11520.4) < 1e-1
    assert abs(grape.getCost([1.0] * N) - 0.397872) < TOL
    group.setGamma(1)
    assert abs(group.getCost([0.0] * M) - 11520.4) < 1e-1
    assert abs(group.getCost(oo.linspace(0, 7, M)) - 48360.2) < 1e-1
    grape.setGamma(0)
    group.setGamma(0)


# ---- tests/ControlBasisTests.cpp ----
def simple_basis():
    N, M = 5, 4
    return oo.ControlBasis([1.0] * N, [1.0] * N, [[2.0] * M for _ in range(N)])


def chopped_basis():
    u0 = [1, 1.1, 1.2, 1.3, 1.4, 1.5, 1.6, 1.7, 1.8, 1.9, 2]
    return oo.build_chopped_sine_basis(u0, 1e-1, 1.0, 5)


def check_control_basis(simple, chopped):
    b = simple
    M, N = b.getM(), b.getN()
    assert np.allclose(b.convertControl([0.0] * M), 1.0, atol=1e-8)                       # :58-68
    u2 = b.convertControl([1.0] * M)
    assert np.allclose(u2, 1 + 2.0 * M, atol=1e-8)                                        # :70-78
    assert np.allclose(b.convertControl([0.0] * M, False), u2, atol=1e-8)                 # :80-85 (cache)
    assert np.allclose(b.convertGradient([0.0] * N), 0.0, atol=1e-8)
    assert np.allclose(b.convertGradient([1.0] * N), 2.0 * N, atol=1e-8)                  # :101-109
    assert np.allclose(b.getControlJacobian(), 2.0, atol=1e-8)                            # :114-128
    assert np.allclose(b.convertHessian(np.zeros((N, N))), 0.0, atol=1e-8)
    assert np.allclose(b.convertHessian(np.ones((N, N))), (N * 2.0) ** 2, atol=1e-8)      # :149-162
    assert np.allclose(b.convertHessian(np.eye(N)), N * 4.0, atol=1e-8)                   # :164-183
    c = chopped
    M, N = c.getM(), c.getN()
    assert np.allclose(c.convertControl([0.0] * M), [1 + i * 0.1 for i in range(N)], atol=1e-6)
    res2 = [1, 4.75688, 4.27768, 1.78131, 1.4, 2.5, 2.32654, 1.45476, 1.8, 2.47919, 2]    # :204
    u2 = c.convertControl([1.0] * M)
    assert np.allclose(u2, res2, atol=5e-6)
    assert np.allclose(c.convertControl([0.0] * M, False), u2, atol=5e-6)
    assert np.allclose(c.convertGradient([0.0] * N), 0.0, atol=5e-6)
    assert np.allclose(c.convertGradient([1.0] * N), [6.31375, 3.58979e-09, 1.96261, 7.17958e-09, 1], atol=5e-6)  # :237
    jac = [[0, 0, 0, 0, 0],
           [0.309017, 0.587785, 0.809017, 0.951057, 1],
           [0.587785, 0.951057, 0.951057, 0.587785, 3.58979e-09],
           [0.809017, 0.951057, 0.309017, -0.587785, -1],
           [0.951057, 0.587785, -0.587785, -0.951057, -7.17959e-09],
           [1, 3.58979e-09, -1, -7.17959e-09, 1],
           [0.951057, -0.587785, -0.587785, 0.951057, 1.07694e-08],
           [0.809017, -0.951057, 0.309017, 0.587785, -1],
           [0.587785, -0.951057, 0.951057, -0.587785, -1.43592e-08],
           [0.309017, -0.587785, 0.809017, -0.951057, 1],
           [0, -0, 0, -0, 0]]                                                              # :250-262
    assert np.allclose(c.getControlJacobian(), jac, atol=5e-6)
    assert np.allclose(c.convertHessian(np.zeros((N, N))), 0.0, atol=1e-10)
    h2 = [[39.8635, 0, 12.3914, 0, 6.3138], [0, 0, 0, 0, 0], [12.3914, 0, 3.8518, 0, 1.9626],
          [0, 0, 0, 0, 0], [6.3138, 0, 1.9626, 0, 1.0]]                                    # :297-303
    assert np.allclose(c.convertHessian(np.ones((N, N))), h2, atol=1e-4)
    H3 = np.ones((N, N))
    idx = 0.0
    for i in range(N):
        for j in range(i, N):
            H3[i, j] = idx
            H3[j, i] = idx
            idx += 0.01
    h3 = [[14.8420, -3.5725, 3.3413, -1.8170, 1.6800], [-3.5725, 1.6547, -0.8321, 0.4766, -0.4938],
          [3.3413, -0.8321, 1.1382, -0.3595, 0.4339], [-1.8170, 0.4766, -0.3595, 0.3759, -0.1662],
          [1.6800, -0.4938, 0.4339, -0.1662, 0.3300]]                                      # :330-335
    assert np.allclose(c.convertHessian(H3), h3, atol=1e-4)


def test_control_basis_goldens():
    check_control_basis(simple_basis(), chopped_basis())


# ---- tests/SiteSetTests.cpp: operator matrix elements (include/BH_sites.h:129-171) ----
def test_operator_matrix_elements():
    D = 6
    op = ob.boson_ops(D)
    for n in range(D):
        assert op["N"][n, n] == n
        assert op["N(N-1)"][n, n] == n * n - n
        assert op["NN"][n, n] == n * n
    for n in range(1, D):
        assert op["A"][n - 1, n] == pytest.approx(np.sqrt(n))
        assert op["Adag"][n, n - 1] == pytest.approx(np.sqrt(n))
    assert np.allclose(op["Adag"] @ op["A"], op["N"])
    # Mott state: <N> = 1 on every site, largest eigenvalue of <adag_i a_j> equals the filling
    psi = ob.product_state([1, 1, 1, 1], D)
    dense = psi.to_dense()
    assert abs(dense[1, 1, 1, 1]) == 1.0
