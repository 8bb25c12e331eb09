"""Re-entrancy of the C ABI: the reference steps one const BH_tDMRG from several std::threads
(src/OptimalControl.cpp:424-430, src/BH_tDMRG.cpp:113-115) and evaluates independent OptimalControl objects concurrently
(tests/GradientTests.cpp:261-285).  Calls on distinct MPS / stores must give the sequential results (tolerance 1e-11;
the engine is deterministic, so the results are in fact identical).  Also the step-graph cache must never replay a
graph captured for an MPS that has been destroyed (buffer addresses are reused by the allocator)."""
import threading

import numpy as np
import pytest

from conftest import golden_state, load_golden, to_host, to_oracle

pytestmark = pytest.mark.gpu


def _setup(name="golden_L8_maxm.npz"):
    import optimalcontrolmps_b200 as oc
    z = load_golden(name)
    L, d, Np, J, cs, ce, T, ts, cutoff, maxm, M, gamma, N = z["params"]
    L, d, N = int(L), int(d), int(N)
    init, target = golden_state(z, "init"), golden_state(z, "target")
    cap = max([int(maxm)] + init.bond_dims() + target.bond_dims())
    st = oc.BH_tDMRG(oc.BoseHubbard(L, d), J, ts, oc.Args("Cutoff=", cutoff, "Maxm=", int(maxm)), chi_cap=cap)
    return oc, z, st, init, target, N, float(gamma)


def _flat(h):
    return np.concatenate([a.ravel() for a in h.A])


def test_two_threads_step_two_mps_on_one_stepper():
    oc, z, st, init, target, N, gamma = _setup()
    us = np.linspace(2.5, 9.0, 13)

    def evolve(start, fwd, out, key):
        dev = st.to_device(to_host(start))
        for k in range(len(us) - 1):
            st.step(dev, us[k], us[k + 1], fwd)
        out[key] = dev.download()

    seq, par = {}, {}
    evolve(init, True, seq, "a")
    evolve(target, False, seq, "b")
    for rep in range(3):                      # repeated: the second and third pass replay captured step graphs
        ths = [threading.Thread(target=evolve, args=(init, True, par, "a")),
               threading.Thread(target=evolve, args=(target, False, par, "b"))]
        for t in ths:
            t.start()
        for t in ths:
            t.join()
        for k in ("a", "b"):
            assert par[k].bond_dims() == seq[k].bond_dims()
            assert np.max(np.abs(_flat(par[k]) - _flat(seq[k]))) < 1e-11


def test_concurrent_cost_and_gradient_on_two_problems():
    oc, z, st, init, target, N, gamma = _setup()
    rng = np.random.default_rng(5)
    ctrls = [list(np.array(z["u"]) + 0.05 * rng.normal(size=N)) for _ in range(2)]
    probs = [oc.OptimalControl(to_host(target), to_host(init), st, N, gamma) for _ in range(2)]
    for p in probs:
        p.setThreadCount(2)
    seq = [(p.getCost(c), p.getAnalyticGradient(c, True)) for p, c in zip(probs, ctrls)]
    par = [None, None]

    def run(k):
        par[k] = (probs[k].getCost(ctrls[k]), probs[k].getAnalyticGradient(ctrls[k], True))

    ths = [threading.Thread(target=run, args=(k,)) for k in range(2)]
    for t in ths:
        t.start()
    for t in ths:
        t.join()
    for k in range(2):
        assert abs(par[k][0] - seq[k][0]) < 1e-11
        assert np.max(np.abs(np.array(par[k][1]) - np.array(seq[k][1]))) < 1e-11


def test_step_graphs_do_not_outlive_their_mps():
    """An MPS is stepped until its step graph is captured and replayed, then destroyed; new MPS (which may land on the
    same addresses) of a DIFFERENT state must evolve exactly as they do in a fresh sequence."""
    oc, z, st, init, target, N, gamma = _setup()
    us = np.linspace(3.0, 6.0, 6)
    from oracle import bh_mps as ob

    def evolve(start, fwd):
        dev = st.to_device(to_host(start))
        for k in range(len(us) - 1):
            st.step(dev, us[k], us[k + 1], fwd)
        res = dev.download()
        del dev
        return res

    ref_a, ref_b = evolve(init, True), evolve(target, False)
    for rep in range(4):
        a = evolve(init, True)
        b = evolve(target, False)
        assert a.bond_dims() == ref_a.bond_dims() and b.bond_dims() == ref_b.bond_dims()
        assert np.max(np.abs(_flat(a) - _flat(ref_a))) < 1e-11
        assert np.max(np.abs(_flat(b) - _flat(ref_b))) < 1e-11
    want = to_oracle(to_host(init))
    so = ob.BHStepper(st.L, st.D, st.J, st.getTstep(), ob.TruncArgs(cutoff=st.args.getReal("Cutoff"), maxm=st.args.getInt("Maxm")))
    for k in range(len(us) - 1):
        so.step(want, us[k], us[k + 1], True)
    assert abs(abs(ob.overlap(want, to_oracle(ref_a))) - 1.0) < 1e-11


def test_larger_charges_after_graph_capture_are_not_truncated_away():
    """The block-kernel grid is sized from the largest charge uploaded so far and is baked into captured graphs: a state
    with more bosons uploaded later must get graphs of its own (or a loud error), never silently skipped blocks."""
    import optimalcontrolmps_b200 as oc
    from conftest import random_symmetric_mps
    from oracle import bh_mps as ob
    L, d, chi = 6, 4, 24
    D = d + 1
    ctx = oc.Context(0)                       # a context of its own: qmax starts at 0
    st = oc.BH_tDMRG(oc.BoseHubbard(L, d), 1.0, 0.01, oc.Args("Cutoff=", 1e-10, "Maxm=", chi), chi_cap=chi, ctx=ctx)
    so = ob.BHStepper(L, D, 1.0, 0.01, ob.TruncArgs(cutoff=1e-10, maxm=chi))
    for Npart in (3, 12):
        psi = random_symmetric_mps(L, D, Npart, chi, seed=Npart)
        dev = st.to_device(to_host(psi))
        want = psi.copy()
        for k in range(4):
            st.step(dev, 3.0, 3.5, True)
            so.step(want, 3.0, 3.5, True)
        got = to_oracle(dev.download())
        assert got.bond_dims() == want.bond_dims()
        assert abs(abs(ob.overlap(want, got)) - 1.0) < 1e-10


def test_trim_releases_idle_workspaces_and_everything_still_works():
    """ocmps_ctx_trim frees the pooled per-chain workspaces (including the Hessian buffers); later calls re-create them and give
    the same numbers."""
    oc, z, st, init, target, N, gamma = _setup("golden_L6_maxm.npz")
    u = list(z["u"])
    p = oc.OptimalControl(to_host(target), to_host(init), st, N, gamma)
    H0 = np.array(p.getHessian(u, True))
    c0 = p.getCost(u, False)
    st.ctx.trim()
    st.ctx.trim()                               # idempotent
    H1 = np.array(p.getHessian(u, True))
    assert np.array_equal(H0, H1)
    assert p.getCost(u, True) == c0
    assert np.max(np.abs(H0 - z["hessian"])) / np.max(np.abs(z["hessian"])) < 1e-6


def test_state_in_another_gauge_is_moved_to_site_1():
    """A host state whose orthogonality limits are not (0, 2) -- here: nothing assumed orthogonal, and centre at the last site -- is
    gauged to site 1 with the engine's own moves on upload (psi.position(), src/BH_tDMRG.cpp:139-148) and then evolves like the
    oracle's copy of it."""
    import optimalcontrolmps_b200 as oc
    from conftest import random_symmetric_mps
    from oracle import bh_mps as ob
    L, d, chi = 6, 4, 16
    D = d + 1
    st = oc.BH_tDMRG(oc.BoseHubbard(L, d), 1.0, 0.01, oc.Args("Cutoff=", 1e-10, "Maxm=", chi), chi_cap=chi)
    so = ob.BHStepper(L, D, 1.0, 0.01, ob.TruncArgs(cutoff=1e-10, maxm=chi))
    base = random_symmetric_mps(L, D, 6, chi, seed=5)
    for case in ("none", "right_end"):
        psi = base.copy()
        if case == "none":                              # destroy the canonical form: rescale and mix nothing -> claim no orthogonality
            psi.A[2] = psi.A[2] * 1.7
            psi.llim, psi.rlim = 0, L + 1
        else:
            psi.position(L)                             # centre at the last site
        want = psi.copy()
        want.position(1)
        dev = st.to_device(to_host(psi))
        got = to_oracle(dev.download())
        assert (got.llim, got.rlim) == (0, 2)
        assert abs(ob.overlap(want, got) - ob.overlap(want, want)) < 1e-11 * abs(ob.overlap(want, want))
        assert abs(dev.norm() - want.norm()) < 1e-11 * want.norm()
        for k in range(2):
            st.step(dev, 3.0, 3.5, True)
            so.step(want, 3.0, 3.5, True)
        got = to_oracle(dev.download())
        assert got.bond_dims() == want.bond_dims()
        assert abs(abs(ob.overlap(want, got)) - 1.0) < 1e-10
