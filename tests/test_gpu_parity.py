"""GPU parity tests: libocmps (through the C ABI via ctypes) against the oracle on the same seeded inputs
and against the committed golden vectors.  Tolerances are the ones BASELINE.json's north_star states:
cost/fidelity 1e-9 relative, gradient 1e-7, Hessian 1e-6, retained bond dimensions identical."""
import itertools

import numpy as np
import pytest

from conftest import golden_state, load_golden, to_host, to_oracle

pytestmark = pytest.mark.gpu

TOL_COST, TOL_GRAD, TOL_HESS = 1e-9, 1e-7, 1e-6


def rel(a, b):
    a, b = np.asarray(a), np.asarray(b)
    return float(np.max(np.abs(a - b)) / max(np.max(np.abs(b)), 1e-300))


def make_stepper(oc, L, d, J, ts, cutoff, maxm, cap_states=()):
    a = oc.Args("Cutoff=", cutoff) if maxm is None else oc.Args("Cutoff=", cutoff, "Maxm=", maxm)
    cap = None
    if maxm is not None:
        cap = max([maxm] + [max(s.bond_dims()) for s in cap_states])
    return oc.BH_tDMRG(oc.BoseHubbard(L, d), J, ts, a, chi_cap=cap)


GOLDENS = ["golden_L3.npz", "golden_L5.npz", "golden_L6_maxm.npz", "golden_L8_maxm.npz"]


@pytest.fixture(scope="module", params=GOLDENS)
def golden(request):
    import optimalcontrolmps_b200 as oc
    z = load_golden(request.param)
    L, d, Np, J, cs, ce, T, ts, cutoff, maxm, M, gamma, N = z["params"]
    L, d, Np, M, N = int(L), int(d), int(Np), int(M), int(N)
    maxm = None if maxm < 0 else int(maxm)
    init, target = golden_state(z, "init"), golden_state(z, "target")
    st = make_stepper(oc, L, d, J, ts, cutoff, maxm, (init, target))
    return dict(oc=oc, z=z, L=L, d=d, J=J, cs=cs, ce=ce, T=T, ts=ts, cutoff=cutoff, maxm=maxm, M=M, gamma=gamma, N=N,
                init=init, target=target, st=st)


def test_single_steps_match_golden(golden):
    g = golden
    oc, z, st = g["oc"], g["z"], g["st"]
    from oracle import bh_mps as ob
    for name, start, a, b, fwd in [("step_fwd", g["init"], 3.0, 4.5, True), ("step_bwd", g["target"], 7.0, 6.0, False)]:
        dev = st.to_device(to_host(start))
        st.step(dev, a, b, fwd)
        got = to_oracle(dev.download())
        want = golden_state(z, name)
        assert got.bond_dims() == want.bond_dims()
        assert got.check_charges() == 0.0
        assert abs(abs(ob.overlap(want, got)) - 1.0) < 1e-12
        assert abs(dev.norm() - 1.0) < 1e-13
        # host-IQMPS flavour of step() (drop-in behaviour of BH_tDMRG::step) gives the same state
        h = to_host(start)
        st.step(h, a, b, fwd)
        assert abs(abs(ob.overlap(want, to_oracle(h))) - 1.0) < 1e-12


def test_grape_cost_gradient_hessian_match_golden(golden):
    g = golden
    oc, z, st, N = g["oc"], g["z"], g["st"], g["N"]
    ocg = oc.OptimalControl(to_host(g["target"]), to_host(g["init"]), st, N, g["gamma"])
    u = list(z["u"])
    cost = ocg.getCost(u)
    assert abs(cost - float(z["cost"])) / abs(float(z["cost"])) < TOL_COST
    fid = ocg.getFidelityForAllT(u, False)
    assert rel(fid, z["fidelities"]) < TOL_COST
    assert np.array_equal(ocg.psi_t.bond_dims(), z["psi_dims"])          # identical retained bond dimensions
    grad = ocg.getAnalyticGradient(u, True)
    assert rel(grad, z["grad"]) < TOL_GRAD
    assert np.array_equal(ocg.xi_t.bond_dims(), z["xi_dims"])
    assert rel(ocg.divT, z["divT"]) < TOL_GRAD
    H = np.array(ocg.getHessian(u, False))
    assert rel(H, z["hessian"]) < TOL_HESS
    kp = st.exactApplyMPO(ocg.psi_t.get(N // 2))
    assert kp.bond_dims() == z["applyK_dims"].tolist()
    assert abs(kp.norm() - float(z["applyK_norm"])) < 1e-10 * float(z["applyK_norm"])
    ov = oc.overlapC(ocg.psi_t.get(N // 2), kp)
    assert abs(ov - complex(z["applyK_ovl"])) < 1e-10 * abs(complex(z["applyK_ovl"]))


def test_group_cost_gradient_hessian_match_golden(golden):
    g = golden
    oc, z, st, N, M = g["oc"], g["z"], g["st"], g["N"], g["M"]
    u0 = oc.SeedGenerator.linspace(g["cs"], g["ce"], N)
    basis = oc.ControlBasisFactory.buildChoppedSineBasis(u0, g["ts"], g["T"], M)
    ocg = oc.OptimalControl(to_host(g["target"]), to_host(g["init"]), st, basis, g["gamma"])
    c = list(z["c"])
    assert abs(ocg.getCost(c) - float(z["group_cost"])) / abs(float(z["group_cost"])) < TOL_COST
    assert rel(ocg.getAnalyticGradient(c, True), z["group_grad"]) < TOL_GRAD
    assert rel(ocg.getHessian(c, False), z["group_hessian"]) < TOL_HESS


def test_threads_bfgs_and_determinism(golden):
    """tests/GradientTests.cpp:261-285, tests/HessianTests.cpp:254-269: sequential vs threaded agree to 1e-11;
    here threadCount selects the number of concurrent CUDA streams / rows in flight."""
    g = golden
    oc, z, st, N = g["oc"], g["z"], g["st"], g["N"]
    u = list(z["u"])
    a = oc.OptimalControl(to_host(g["target"]), to_host(g["init"]), st, N, g["gamma"])
    g1 = np.array(a.getAnalyticGradient(u))
    H1 = np.array(a.getHessian(u, False))
    b = oc.OptimalControl(to_host(g["target"]), to_host(g["init"]), st, N, g["gamma"])
    b.setThreadCount(4)
    g4 = np.array(b.getAnalyticGradient(u))
    H4 = np.array(b.getHessian(u, False))
    assert np.max(np.abs(g1 - g4)) < 1e-11
    assert np.max(np.abs(H1 - H4)) < 1e-11
    with pytest.raises(ValueError):
        b.setThreadCount(0)                                     # src/OptimalControl.cpp:56
    c = oc.OptimalControl(to_host(g["target"]), to_host(g["init"]), st, N, g["gamma"], True)
    assert np.max(np.abs(np.array(c.getAnalyticGradient(u)) - g1)) < 1e-11   # BFGS branch (:217-229)
    assert c.useBFGS()


def test_overlaps_against_oracle(golden):
    g = golden
    oc, st = g["oc"], g["st"]
    from oracle import bh_mps as ob
    a, b = st.to_device(to_host(g["init"])), st.to_device(to_host(g["target"]))
    assert abs(oc.overlapC(a, b) - ob.overlap(g["init"], g["target"])) < 1e-13
    assert abs(oc.overlapC_K(a, b) - ob.overlap_K(g["init"], g["target"])) < 1e-12
    assert abs(oc.overlapC(b, a) - np.conj(ob.overlap(g["init"], g["target"]))) < 1e-13
    assert abs(oc.overlapC(a, a) - 1.0) < 1e-13


def test_live_oracle_parity_random_controls():
    """Same seeded inputs through the oracle and the GPU (not via fixtures), including a capped Maxm."""
    import optimalcontrolmps_b200 as oc
    from oracle import bh_mps as ob, optimal_control as oo, ground_state as og
    L, d, Np, J, ts = 6, 3, 6, 1.0, 1e-2
    D = d + 1
    N = 9
    psi_i = og.ground_state_ed(L, D, Np, J, 2.5)
    psi_f = og.ground_state_ed(L, D, Np, J, 30.0)
    for cutoff, maxm in [(1e-8, None), (1e-10, 10), (1e-6, 6)]:
        st_o = ob.BHStepper(L, D, J, ts, ob.TruncArgs(cutoff=cutoff, maxm=maxm))
        st = make_stepper(oc, L, d, J, ts, cutoff, maxm, (psi_i, psi_f))
        u = list(np.random.default_rng(100).uniform(2, 40, N))
        o = oo.OptimalControl(psi_f, psi_i, st_o, N=N, gamma=1e-6)
        gq = oc.OptimalControl(to_host(psi_f), to_host(psi_i), st, N, 1e-6)
        co, cg = o.getCost(u), gq.getCost(u)
        assert abs(co - cg) / abs(co) < TOL_COST
        assert np.array_equal(gq.psi_t.bond_dims(), np.array([p.bond_dims() for p in o.psi_t]))
        assert rel(gq.getAnalyticGradient(u), o.getAnalyticGradient(u)) < TOL_GRAD
        assert rel(gq.getHessian(u, False), o.getHessian(u, False)) < TOL_HESS


def test_new_control_cache_contract_on_gpu():
    """tests/SequencingTest.cpp on the GPU path: L=3, Npart=3, d=3, J=2, Cutoff 1e-7."""
    import optimalcontrolmps_b200 as oc
    from oracle import ground_state as og
    L, d, J, T, ts = 3, 3, 2.0, 0.2, 1e-2
    N = int(T / ts + 1)
    psi_i = og.ground_state_ed(L, d + 1, 3, J, 2.0)
    psi_f = og.ground_state_ed(L, d + 1, 3, J, 12.0)
    st = make_stepper(oc, L, d, J, ts, 1e-7, None)
    rng = np.random.default_rng(3)
    u = list(rng.uniform(5, 15, N))
    ref = oc.OptimalControl(to_host(psi_f), to_host(psi_i), st, N, 0)
    c0, g0, H0 = ref.getCost(u, True), np.array(ref.getAnalyticGradient(u, True)), np.array(ref.getHessian(u, True))
    calls = {"c": lambda o, nc: o.getCost(u, nc), "g": lambda o, nc: np.array(o.getAnalyticGradient(u, nc)),
             "h": lambda o, nc: np.array(o.getHessian(u, nc))}
    for order in itertools.permutations("cgh"):                # :116-198
        o = oc.OptimalControl(to_host(psi_f), to_host(psi_i), st, N, 0)
        res = {}
        for k, name in enumerate(order):
            res[name] = calls[name](o, k == 0)
        assert abs(res["c"] - c0) < 1e-10
        assert np.max(np.abs(res["g"] - g0)) < 1e-10
        assert np.max(np.abs(res["h"] - H0)) < 1e-10
    ref.setBFGS(True)                                           # :127-133
    assert abs(ref.getCost(u, True) - c0) < 1e-10
    assert np.max(np.abs(np.array(ref.getAnalyticGradient(u, False)) - g0)) < 1e-10
    ref.setBFGS(False)
    ref.getAnalyticGradient(u, True)
    u2 = list(rng.uniform(2, 20, N))
    assert abs(ref.getCost(u2, False) - c0) < 1e-10            # stale cost (:238-246)
    assert np.max(np.abs(np.array(ref.getAnalyticGradient(u2, False)) - g0)) < 1e-10   # stale gradient (:248-256)
    assert abs(ref.getCost(u2, True) - c0) > 1e-10
    ref.getAnalyticGradient(u, True)
    u3 = list(rng.uniform(1, 4, N))
    Hf = np.array(ref.getHessian(u3, False))
    assert np.max(np.abs(Hf - H0)) > 1e-10                     # rows re-propagated with the passed control (:258-266)
    Ht = np.array(ref.getHessian(u3, True))
    assert np.max(np.abs(Ht - Hf)) > 1e-10


def test_edge_cases():
    import optimalcontrolmps_b200 as oc
    from oracle import bh_mps as ob, ground_state as og
    from optimalcontrolmps_b200 import _lib
    # L = 2 (one gate), odd and even L with product states (bond dimension 1 everywhere)
    for L in (2, 3, 4):
        d = 2
        st_o = ob.BHStepper(L, d + 1, 1.0, 1e-2, ob.TruncArgs(cutoff=1e-10))
        st = make_stepper(oc, L, d, 1.0, 1e-2, 1e-10, None)
        po = ob.product_state([1] * L, d + 1)
        dev = st.to_device(to_host(po))
        for k in range(3):
            st_o.step(po, 2.0 + k, 3.0 + k, True)
            st.step(dev, 2.0 + k, 3.0 + k, True)
        got = to_oracle(dev.download())
        assert got.bond_dims() == po.bond_dims()
        assert abs(abs(ob.overlap(po, got)) - 1.0) < 1e-12
    # capacity overflow is reported, not silently truncated
    psi = og.ground_state_ed(6, 4, 6, 1.0, 2.5)
    st = oc.BH_tDMRG(oc.BoseHubbard(6, 3), 1.0, 1e-2, oc.Args("Cutoff=", 1e-12), chi_cap=8)
    with pytest.raises(_lib.OcmpsError):
        st.to_device(to_host(psi))
    # schedule mirrors the oracle's gate order
    for L in (2, 3, 4, 5, 8):
        st = make_stepper(oc, L, 2, 1.0, 1e-2, 1e-10, None)
        gates = [(op[1], op[1] + 1) for op in st.schedule() if op[0] == 1]
        assert gates == ob.gate_order(L)


def test_batched_controls_match_single_evaluations(golden):
    """ocmps_sweep_batch: several independent controls in flight on one GPU give exactly the single-evaluation results."""
    g = golden
    oc, z, st, N, M = g["oc"], g["z"], g["st"], g["N"], g["M"]
    u0 = oc.SeedGenerator.linspace(g["cs"], g["ce"], N)
    rng = np.random.default_rng(77)
    probs, ctrls = [], []
    for k in range(3):
        basis = oc.ControlBasisFactory.buildChoppedSineBasis(u0, g["ts"], g["T"], M)
        probs.append(oc.OptimalControl(to_host(g["target"]), to_host(g["init"]), st, basis, g["gamma"]))
        ctrls.append(list(rng.uniform(-1, 1, M)))
    ctrls[0] = list(z["c"])
    res = oc.batch_cost_gradient(probs, ctrls)
    assert abs(res[0][0] - float(z["group_cost"])) / abs(float(z["group_cost"])) < TOL_COST
    assert rel(res[0][1], z["group_grad"]) < TOL_GRAD
    for k in range(3):
        single = oc.OptimalControl(to_host(g["target"]), to_host(g["init"]), st, probs[k].basis, g["gamma"])
        gs = single.getAnalyticGradient(ctrls[k], True)
        cs = single.getCost(ctrls[k], False)
        assert abs(cs - res[k][0]) < 1e-12
        assert np.max(np.abs(np.array(gs) - np.array(res[k][1]))) < 1e-12
        assert abs(probs[k].getCost(ctrls[k], False) - cs) < 1e-12       # the problem is left in the cached state


@pytest.mark.parametrize("L,d,Np,chi", [(6, 1, 3, 8), (6, 2, 6, 16), (7, 3, 7, 24), (5, 6, 6, 30), (4, 7, 9, 40), (10, 4, 10, 70)])
def test_steps_match_oracle_for_every_local_dimension(L, d, Np, chi):
    """The kernels are specialised at compile time on the local dimension D = d+1 (2..8) and on the block shapes; walk
    through all of them: two forward and one backward Trotter step of a random number-conserving MPS against the oracle
    (identical bond dimensions, same state to 1e-9, charges conserved)."""
    import optimalcontrolmps_b200 as oc
    from oracle import bh_mps as ob
    from conftest import random_symmetric_mps
    psi = random_symmetric_mps(L, d + 1, Np, chi, seed=100 * L + d)
    cutoff = 1e-9
    so = ob.BHStepper(L, d + 1, 1.0, 2e-2, ob.TruncArgs(cutoff=cutoff, maxm=chi))
    st = oc.BH_tDMRG(oc.BoseHubbard(L, d), 1.0, 2e-2, oc.Args("Cutoff=", cutoff, "Maxm=", chi),
                     chi_cap=max(chi, max(psi.bond_dims())))
    dev = st.to_device(to_host(psi))
    po = psi.copy()
    for (a, b, fwd) in [(2.0, 3.0, True), (3.0, 5.0, True), (5.0, 4.0, False)]:
        so.step(po, a, b, fwd)
        st.step(dev, a, b, fwd)
        got = to_oracle(dev.download())
        assert got.bond_dims() == po.bond_dims()
        assert abs(abs(ob.overlap(po, got)) - 1.0) < 1e-9
        assert abs(dev.norm() - 1.0) < 1e-12
        assert got.check_charges() == 0.0


def test_site_expectation_values_on_resident_slices(golden):
    """Observables on the HBM-resident slices (include/correlations.hpp:99-117) against the oracle: <N_j>, <N_j(N_j-1)>,
    <N_j^2> of every slice of a forward sweep; the particle number is conserved; the built-in norm check holds."""
    g = golden
    oc, z, st, N = g["oc"], g["z"], g["st"], g["N"]
    from oracle import observables as obs
    ocg = oc.OptimalControl(to_host(g["target"]), to_host(g["init"]), st, N, g["gamma"])
    u = list(z["u"])
    ocg.getCost(u)                                        # forward sweep: psi_t resident
    vals, nrm = ocg.psi_t.expectationValues(("N", "N(N-1)", "NN"), return_norm=True)
    assert vals.shape == (N, g["L"], 3)
    assert np.max(np.abs(nrm - 1.0)) < 1e-12              # centre at site 1 in every slice, unit norm
    npart = float(np.sum(obs.expectation_values(g["init"], np.arange(g["d"] + 1, dtype=float))))
    assert np.max(np.abs(vals[:, :, 0].sum(axis=1) - npart)) < 1e-10
    n = np.arange(g["d"] + 1, dtype=float)
    for i in sorted(set([0, 1, N // 3, N // 2, N - 1])):
        po = to_oracle(ocg.psi_t.get(i).download())
        for k, diag in enumerate((n, n * (n - 1.0), n * n)):
            want = obs.expectation_values(po, diag)
            assert np.max(np.abs(vals[i, :, k] - want)) < 1e-11
    assert np.max(np.abs(vals[:, :, 2] - vals[:, :, 1] - vals[:, :, 0])) < 1e-11      # NN = N(N-1) + N
    sub = ocg.psi_t.expectationValues(("N",), first=2, count=3)
    assert np.array_equal(sub[:, :, 0], vals[2:5, :, 0])


def test_entanglement_entropy_on_resident_slices(golden):
    """Bond entropies of the resident slices (include/correlations.hpp:119-148) against the oracle; the store is unchanged."""
    g = golden
    oc, z, st, N = g["oc"], g["z"], g["st"], g["N"]
    from oracle import observables as obs
    ocg = oc.OptimalControl(to_host(g["target"]), to_host(g["init"]), st, N, g["gamma"])
    ocg.getCost(list(z["u"]))
    dims = np.array(ocg.psi_t.bond_dims())
    S = ocg.psi_t.entanglementEntropy()
    assert S.shape == (N, g["L"] - 1)
    assert np.array_equal(np.array(ocg.psi_t.bond_dims()), dims)
    for i in sorted(set([0, 1, N // 2, N - 1])):
        want = obs.entanglement_entropy(to_oracle(ocg.psi_t.get(i).download()))
        assert np.max(np.abs(S[i] - want)) < 1e-10
    assert np.array_equal(ocg.psi_t.entanglementEntropy(first=1, count=2), S[1:3])


def test_reference_costtests_goldens_on_the_gpu():
    """The golden fidelities and costs the reference's own tests hold (tests/CostTests.cpp:67-133, 136-203: L=5, Npart=5,
    d=5, U 2 -> 50, T=0.1, tstep=0.01, M=5, Cutoff 1e-8) evaluated by the GPU engine.  Tolerance 1e-5: the reference's
    numbers come from inexact DMRG ground states (visible at t=0); ours from exact diagonalisation."""
    import optimalcontrolmps_b200 as oc
    from oracle import ground_state as og
    FID_LIN = [0.214338, 0.214325, 0.215126, 0.217281, 0.221019, 0.22621, 0.232328, 0.238484, 0.243617, 0.246862, 0.24801]
    FID_ONE = [0.214338, 0.214233, 0.213919, 0.213398, 0.212672, 0.211744, 0.210618, 0.2093, 0.207796, 0.206112, 0.204256]
    FID_GRP = [0.214338, 0.21411, 0.216706, 0.222581, 0.229759, 0.23623, 0.242512, 0.249913, 0.256515, 0.259334, 0.259687]
    TOL = 1e-5
    L, Npart, d = 5, 5, 5
    J, cs, ce, T, ts, M = 1.0, 2.0, 50.0, 0.1, 1e-2, 5
    N = int(T / ts + 1)
    psi_i = og.ground_state_ed(L, d + 1, Npart, J, cs)
    psi_f = og.ground_state_ed(L, d + 1, Npart, J, ce)
    st = make_stepper(oc, L, d, J, ts, 1e-8, None)
    grape = oc.OptimalControl(to_host(psi_f), to_host(psi_i), st, N, 0.0)
    u = oc.SeedGenerator.linspace(cs, ce, N)
    assert abs(grape.getCost(u) - 0.375995) < TOL                                        # :78
    assert np.max(np.abs(np.array(grape.getFidelityForAllT(u, False)) - FID_LIN)) < TOL  # :75
    assert abs(grape.getCost([1.0] * N) - 0.397872) < TOL                                # :93
    assert np.max(np.abs(np.array(grape.getFidelityForAllT([1.0] * N, False)) - FID_ONE)) < TOL
    basis = oc.ControlBasisFactory.buildChoppedSineBasis(oc.SeedGenerator.linspace(cs, ce, N), ts, T, M)
    group = oc.OptimalControl(to_host(psi_f), to_host(psi_i), st, basis, 0.0)
    assert abs(group.getCost([0.0] * M) - 0.375995) < TOL                                # :112
    assert np.max(np.abs(np.array(group.getFidelityForAllT([0.0] * M, False)) - FID_LIN)) < TOL
    c2 = oc.SeedGenerator.linspace(0, 7, M)
    assert abs(group.getCost(c2) - 0.370157) < TOL                                       # :127
    assert np.max(np.abs(np.array(group.getFidelityForAllT(c2, False)) - FID_GRP)) < TOL
    grape.setGamma(1)                                                                    # :136-203
    assert abs(grape.getCost(oc.SeedGenerator.linspace(cs, ce, N)) - 11520.4) < 1e-1
    group.setGamma(1)
    assert abs(group.getCost(c2) - 48360.2) < 1e-1
