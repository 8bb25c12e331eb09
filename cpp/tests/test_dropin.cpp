// Drop-in test of the C++ layer: reads a problem + expected results written by tests/test_gpu_cpp_dropin.py
// (binary, little endian) and drives it exactly like the reference's tests / main programs would:
// BoseHubbard -> BH_tDMRG -> OptimalControl (GRAPE and GROUP) -> BH_nlp callbacks.  Exit code 0 = all checks pass.
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <fstream>
#include <iostream>
#include <thread>
#include <vector>

#include "BH_nlp.hpp"
#include "correlations.hpp"
#include "BH_tDMRG.hpp"
#include "ControlBasisFactory.hpp"
#include "InitializeState.hpp"
#include "OptimalControl.hpp"

static std::ifstream in;
static int failures = 0;

template <typename T> T rd() { T v; in.read(reinterpret_cast<char*>(&v), sizeof(T)); return v; }
static std::vector<double> rdvec() { int n = rd<int>(); std::vector<double> v(n); in.read(reinterpret_cast<char*>(v.data()), 8 * n); return v; }
static std::vector<int> rdivec() { int n = rd<int>(); std::vector<int> v(n); in.read(reinterpret_cast<char*>(v.data()), 4 * n); return v; }

static IQMPS rdstate(int L, int D, int cap) {
  std::vector<int> dims = rdivec(), q = rdivec();
  std::vector<double> t = rdvec();
  std::vector<Cplx> tc(t.size() / 2);
  for (size_t i = 0; i < tc.size(); ++i) tc[i] = Cplx(t[2 * i], t[2 * i + 1]);
  return IQMPS(L, D, cap, dims, q, tc);
}

static double relerr(const std::vector<double>& a, const std::vector<double>& b) {
  double num = 0, den = 1e-300;
  for (size_t i = 0; i < a.size(); ++i) { num = std::max(num, std::fabs(a[i] - b[i])); den = std::max(den, std::fabs(b[i])); }
  return num / den;
}
static void check(bool ok, const char* what, double val) {
  printf("%-46s %s (%.3e)\n", what, ok ? "ok" : "FAIL", val);
  if (!ok) ++failures;
}

int main(int argc, char** argv) {
  if (argc < 2) { fprintf(stderr, "usage: test_dropin problem.bin\n"); return 2; }
  in.open(argv[1], std::ios::binary);
  if (!in) { fprintf(stderr, "cannot open %s\n", argv[1]); return 2; }
  const int L = rd<int>(), d = rd<int>(), N = rd<int>(), M = rd<int>(), maxm = rd<int>(), cap = rd<int>();
  const double J = rd<double>(), tstep = rd<double>(), T = rd<double>(), cutoff = rd<double>(), gamma = rd<double>(), cs = rd<double>(),
               ce = rd<double>();
  auto sites = BoseHubbard(L, d);
  IQMPS psi_i = rdstate(L, d + 1, cap), psi_f = rdstate(L, d + 1, cap);
  std::vector<double> u = rdvec(), c = rdvec();
  const double cost_ref = rd<double>();
  std::vector<double> fid_ref = rdvec(), grad_ref = rdvec(), hess_ref = rdvec();
  const double gcost_ref = rd<double>();
  std::vector<double> ggrad_ref = rdvec(), ghess_ref = rdvec();

  auto stepper = maxm > 0 ? BH_tDMRG(sites, J, tstep, {"Cutoff=", cutoff, "Maxm=", maxm}, cap) : BH_tDMRG(sites, J, tstep, {"Cutoff=", cutoff}, cap);
  // ---- GRAPE (reference constructor order: target first, init second) ----
  OptimalControl<BH_tDMRG> OC(psi_f, psi_i, stepper, (size_t)N, gamma);
  const double cost = OC.getCost(u);
  check(std::fabs(cost - cost_ref) / std::fabs(cost_ref) < 1e-9, "GRAPE cost", std::fabs(cost - cost_ref) / std::fabs(cost_ref));
  check(relerr(OC.getFidelityForAllT(u, false), fid_ref) < 1e-9, "GRAPE fidelities (new_control=false)", relerr(OC.getFidelityForAllT(u, false), fid_ref));
  auto grad = OC.getAnalyticGradient(u, true);
  check(relerr(grad, grad_ref) < 1e-7, "GRAPE gradient", relerr(grad, grad_ref));
  auto H = OC.getHessian(u, false);
  std::vector<double> Hflat;
  for (auto& r : H) Hflat.insert(Hflat.end(), r.begin(), r.end());
  check(relerr(Hflat, hess_ref) < 1e-6, "GRAPE Hessian (new_control=false)", relerr(Hflat, hess_ref));
  OC.setThreadCount(4);
  auto grad4 = OC.getAnalyticGradient(u, true);
  check(relerr(grad4, grad) < 1e-11, "gradient, threadCount 4 vs 1", relerr(grad4, grad));
  bool threw = false;
  try { OC.setThreadCount(0); } catch (const std::invalid_argument&) { threw = true; }
  check(threw, "setThreadCount(0) throws std::invalid_argument", 0.0);
  OC.setBFGS(true);
  auto gradb = OC.getAnalyticGradient(u, true);
  check(relerr(gradb, grad) < 1e-11, "gradient, BFGS mode", relerr(gradb, grad));
  check((int)OC.getPsit().size() == N && linkInd(OC.getPsit().back(), 1).m() >= 1, "getPsit / linkInd", (double)N);
  // stale-cache contract (tests/SequencingTest.cpp:238-246)
  std::vector<double> u2(u);
  for (double& x : u2) x += 1.0;
  check(std::fabs(OC.getCost(u2, false) - cost) < 1e-10, "stale cost with new_control=false", std::fabs(OC.getCost(u2, false) - cost));
  // BH_tDMRG::step is a public entry point of its own (main/AnalyzeQuench.cpp:159)
  IQMPS psi = psi_i;
  stepper.step(psi, u[0], u[1], true);
  check(std::fabs(norm(psi) - 1.0) < 1e-12, "BH_tDMRG::step keeps the norm", std::fabs(norm(psi) - 1.0));
  {   // include/correlations.hpp: <N_j> of the evolved state adds up to the particle number; NN = N(N-1) + N
    auto nj = expectationValues(sites, psi, "N"), n2 = expectationValues(sites, psi, "NN"), nn1 = expectationValues(sites, psi, "N(N-1)");
    auto n0 = expectationValues(sites, psi_i, "N");
    double tot = 0.0, tot0 = 0.0, dev = 0.0;
    for (size_t j = 0; j < nj.size(); ++j) {
      tot += nj[j].real(); tot0 += n0[j].real();
      dev = std::max(dev, std::fabs(n2[j].real() - nn1[j].real() - nj[j].real()));
    }
    check(std::fabs(tot - tot0) < 1e-10 && std::fabs(tot0 - std::round(tot0)) < 1e-8, "expectationValues: particle number conserved", std::fabs(tot - tot0));
    check(dev < 1e-11, "expectationValues: <NN> = <N(N-1)> + <N>", dev);
    check(std::abs(expectationValue(sites, psi, "N", 2) - nj[1]) < 1e-14, "expectationValue(site 2)", 0.0);
    {   // two-point functions: trace of <Adag_i A_j> is the particle number, its diagonal equals <N_i>, and the largest eigenvalue
        // of the one-body density matrix lies between the largest occupation and the particle number
      auto rho = correlationMatrix(sites, psi, "Adag", "A");
      double tr = 0.0, dmax = 0.0, herm = 0.0, nmax = 0.0;
      for (size_t i = 0; i < rho.size(); ++i) {
        tr += rho[i][i].real();
        dmax = std::max(dmax, std::fabs(rho[i][i].real() - nj[i].real()));
        nmax = std::max(nmax, nj[i].real());
        for (size_t j = 0; j < rho.size(); ++j) herm = std::max(herm, std::abs(rho[i][j] - std::conj(rho[j][i])));
      }
      check(std::fabs(tr - tot) < 1e-10 && dmax < 1e-11 && herm < 1e-14, "correlationMatrix(Adag, A): trace, diagonal, hermiticity", std::fabs(tr - tot));
      const double lam = correlationTerm(sites, psi, "Adag", "A");
      check(lam >= nmax - 1e-10 && lam <= tot + 1e-10, "correlationTerm(Adag, A) within [max n_i, N]", lam);
      const Cplx c13 = correlationFunction(sites, psi, "Adag", 1, "A", 3);
      check(std::abs(c13 - rho[0][2]) < 1e-14, "correlationFunction(Adag,1,A,3) = rho[1][3]", std::abs(c13));
      const Cplx aad = correlationFunction(sites, psi, "A", 2, "Adag", 2), ada = correlationFunction(sites, psi, "Adag", 2, "A", 2);
      const double comm = (aad - ada).real();        // [A, Adag] = 1 below the occupation cut-off, -d on the top level
      check(comm <= 1.0 + 1e-12 && comm > 0.9 && std::fabs((aad - ada).imag()) < 1e-13, "same site: <A Adag> - <Adag A> = 1 - (d+1) P(n=d)", comm);
    }
    auto SvN = entanglementEntropy(sites, psi);
    bool sok = (int)SvN.size() == L - 1;
    for (double x : SvN) sok = sok && x > -1e-12 && x < std::log((double)cap) + 1e-9;
    check(sok, "entanglementEntropy: L-1 values within [0, ln chi]", SvN.empty() ? 0.0 : SvN[SvN.size() / 2]);
  }
  {   // re-entrancy: two std::threads step two states on ONE const stepper (src/OptimalControl.cpp:424-430,
      // src/BH_tDMRG.cpp:113-115) and two OptimalControl objects are evaluated concurrently (tests/GradientTests.cpp:261-285)
    const BH_tDMRG& cst = stepper;
    auto run = [&](IQMPS& s, bool fwd) { for (int k = 0; k + 1 < std::min(N, 6); ++k) cst.step(s, fwd ? u[k] : u[N - 1 - k], fwd ? u[k + 1] : u[N - 2 - k], fwd); };
    IQMPS a_seq = psi_i, b_seq = psi_f, a_par = psi_i, b_par = psi_f;
    run(a_seq, true);
    run(b_seq, false);
    std::thread t1([&] { run(a_par, true); }), t2([&] { run(b_par, false); });
    t1.join(); t2.join();
    const double da = std::abs(overlapC(a_seq, a_par)) - 1.0, db = std::abs(overlapC(b_seq, b_par)) - 1.0;
    check(std::fabs(da) < 1e-11 && std::fabs(db) < 1e-11 && a_seq.bondDims() == a_par.bondDims() && b_seq.bondDims() == b_par.bondDims(),
          "two threads on one const stepper = sequential", std::max(std::fabs(da), std::fabs(db)));
    OptimalControl<BH_tDMRG> P1(psi_f, psi_i, stepper, (size_t)N, gamma), P2(psi_f, psi_i, stepper, (size_t)N, gamma);
    std::vector<double> ua(u), ub(u);
    for (size_t i = 0; i < ub.size(); ++i) ub[i] += 0.01 * (double)(i % 3);
    const double c1 = P1.getCost(ua), c2 = P2.getCost(ub);
    double p1 = 0, p2 = 0;
    std::thread t3([&] { p1 = P1.getCost(ua); }), t4([&] { p2 = P2.getCost(ub); });
    t3.join(); t4.join();
    check(std::fabs(p1 - c1) < 1e-11 && std::fabs(p2 - c2) < 1e-11, "concurrent getCost on two problems = sequential", std::max(std::fabs(p1 - c1), std::fabs(p2 - c2)));
  }
  {   // a temporary stepper that has been destroyed must never be handed to exactApplyMPO (registry of live steppers)
    { BH_tDMRG tmp(sites, J, tstep, {"Cutoff=", cutoff, "Maxm=", maxm > 0 ? maxm : cap}, cap); (void)tmp; }
    IQMPS k0 = exactApplyMPO(IQMPO(IQMPO::PropagatorDerivative), psi, stepper.getArgs());     // MPO without an owner: registry lookup
    IQMPS k1 = exactApplyMPO(stepper.propagatorDeriv(u[0]), psi, stepper.getArgs());
    check(std::fabs(std::abs(overlapC(k0, k1)) - norm(k0) * norm(k1)) < 1e-10 * norm(k0) * norm(k1), "exactApplyMPO resolves the live stepper", norm(k0));
  }
  IQMPS kpsi = exactApplyMPO(stepper.propagatorDeriv(u[0]), psi, stepper.getArgs());
  const Cplx k1 = overlapC(psi, kpsi), k2 = overlapC(psi, stepper.propagatorDeriv(u[0]), psi);
  check(std::abs(k1 - k2) < 5e-3 * std::abs(k2), "<psi|K psi> ~ <psi|K|psi> (up to the Maxm truncation)", std::abs(k1 - k2));

  {   // InitializeState (include/InitializeState.hpp): Mott-side ground state on the device; N bosons, energy below the product state's
    IQMPS gs = InitializeState(sites, L, J, ce, maxm > 0 ? maxm : cap, 1E-9);
    auto n = expectationValues(sites, gs, "N");
    double totn = 0.0;
    for (auto& x : n) totn += x.real();
    double hop = 0.0;
    for (int i = 1; i < L; ++i) hop += correlationFunction(sites, gs, "Adag", i, "A", i + 1).real();
    check(std::fabs(totn - L) < 1e-9 && std::fabs(norm(gs) - 1.0) < 1e-12 && hop > 0.0, "InitializeState: particle number, norm, kinetic energy", hop);
  }
  // ---- GROUP ----
  auto u0 = SeedGenerator::linspace(cs, ce, N);
  auto basis = ControlBasisFactory::buildChoppedSineBasis(u0, tstep, T, M);
  OptimalControl<BH_tDMRG> OCG(psi_f, psi_i, stepper, basis, gamma);
  const double gcost = OCG.getCost(c);
  check(std::fabs(gcost - gcost_ref) / std::fabs(gcost_ref) < 1e-9, "GROUP cost", std::fabs(gcost - gcost_ref) / std::fabs(gcost_ref));
  check(relerr(OCG.getAnalyticGradient(c, true), ggrad_ref) < 1e-7, "GROUP gradient", relerr(OCG.getAnalyticGradient(c, false), ggrad_ref));
  auto GH = OCG.getHessian(c, false);
  std::vector<double> GHflat;
  for (auto& r : GH) GHflat.insert(GHflat.end(), r.begin(), r.end());
  check(relerr(GHflat, ghess_ref) < 1e-6, "GROUP Hessian", relerr(GHflat, ghess_ref));

  // ---- BH_nlp callbacks as IPOPT would issue them ----
  BH_nlp nlp(OCG, false);
  Ipopt::Index n, m, nj, nh;
  TNLP::IndexStyleEnum style;
  nlp.get_nlp_info(n, m, nj, nh, style);
  check(n == M && m == N && nj == N * M && nh == (M * M + M) / 2 && style == TNLP::C_STYLE, "BH_nlp::get_nlp_info", (double)n);
  std::vector<double> xl(n), xu(n), gl(m), gu(m), x(c), g(m), gf(n), hv(nh);
  nlp.get_bounds_info(n, xl.data(), xu.data(), m, gl.data(), gu.data());
  check(xl[0] == -20 && xu[0] == 20 && gl[0] == 2.0 && gu[0] == 100, "BH_nlp::get_bounds_info", xl[0]);
  double f = 0;
  nlp.eval_g(n, x.data(), true, m, g.data());
  nlp.eval_f(n, x.data(), false, f);
  nlp.eval_grad_f(n, x.data(), false, gf.data());
  check(std::fabs(f - gcost_ref) / std::fabs(gcost_ref) < 1e-9, "BH_nlp::eval_f after eval_g(new_x)", std::fabs(f - gcost_ref));
  check(relerr(gf, ggrad_ref) < 1e-7, "BH_nlp::eval_grad_f (new_x=false)", relerr(gf, ggrad_ref));
  nlp.eval_h(n, x.data(), false, 2.0, m, nullptr, false, nh, nullptr, nullptr, hv.data());
  check(std::fabs(hv[0] - 2.0 * ghess_ref[0]) < 1e-6 * std::fabs(2.0 * ghess_ref[0]) + 1e-12, "BH_nlp::eval_h (obj_factor 2)", hv[0]);

  printf("%s: %d failure(s)\n", failures ? "FAILED" : "PASSED", failures);
  return failures ? 1 : 0;
}
