// Derivative-free ramp optimisation (the reference's main/AmoebaOpt.cpp): Nelder-Mead over the GROUP coefficients with a quadratic
// penalty that keeps the control inside [2, 100] (OCWrapper, main/AmoebaOpt.cpp:13-52).  Same input file and keys as OptimizeRamp
// plus gammaBound (100) and, new here, parallelEvals (4): the independent evaluations of the simplex construction and of a shrink
// are dealt to that many copies of the problem, each evaluated by its own host thread -- the C ABI is re-entrant, the copies run
// concurrently on one GPU.   usage: AmoebaOpt InputFile_BHcontrol [seed]
// Output: AmoebaResult.txt (best cost, coefficients), BHrampInitialFinal.txt (time, initial control, initial fidelity, final control,
// final fidelity), AmoebaHistory.txt (iteration, best cost, function evaluations).
#include "Amoeba.hpp"
#include "BH_tDMRG.hpp"
#include "ControlBasisFactory.hpp"
#include "InitializeState.hpp"
#include "OptimalControl.hpp"
#include "SeedGenerator.hpp"

#include <cstdio>
#include <ctime>
#include <fstream>
#include <memory>
#include <thread>

using namespace itensor;
using OC_t = OptimalControl<BH_tDMRG>;

// cost + gammaBound * sum of squared bound violations of the control (main/AmoebaOpt.cpp:19-35)
struct PenalisedCost {
  OC_t& oc;
  double uMin, uMax, gammaBound;
  double penalty(const std::vector<double>& c) const {
    double acc = 0.0;
    for (double u : oc.getControl(c)) {
      if (u > uMax) acc += (u - uMax) * (u - uMax);
      if (u < uMin) acc += (u - uMin) * (u - uMin);
    }
    return gammaBound * acc;
  }
  double operator()(const std::valarray<double>& x) const {
    const std::vector<double> c(std::begin(x), std::end(x));
    return oc.getCost(c) + penalty(c);
  }
};

int main(int argc, char* argv[]) {
  if (argc < 2) { std::printf("Usage: %s InputFile_BHcontrol [seed]\n", argv[0]); return 0; }
  const InputGroup input(argv[1], "input");
  const double tstep = input.getReal("tstep", 1e-2), T = input.getReal("T");
  const int N = input.getInt("N"), Npart = input.getInt("Npart"), locDim = input.getInt("d"), M = input.getInt("M");
  const double gamma = input.getReal("gamma", 0), threshold = input.getReal("threshold", 1e-7), gammaBound = input.getReal("gammaBound", 100);
  const double optTol = input.getReal("optTol", 1e-7);
  const int maxBondDim = input.getInt("maxBondDim", 100), maxFun = input.getInt("maxFun", 5000), workers = std::max(1, input.getInt("parallelEvals", 4));
  const double J = 1.0, U_i = 2.5, U_f = 50;
  const int seed = argc > 2 ? std::stoi(argv[2]) : 1;
  std::srand((unsigned)seed * (unsigned)std::time(nullptr));     // main/AmoebaOpt.cpp:100

  auto sites = BoseHubbard(N, locDim);
  auto u0 = SeedGenerator::linsigmoidSeed(U_i, U_f, (size_t)(T / tstep + 1));
  auto psi_i = InitializeState(sites, Npart, J, u0.front(), maxBondDim, threshold);
  auto psi_f = InitializeState(sites, Npart, J, u0.back(), maxBondDim, threshold);
  auto stepper = BH_tDMRG(sites, J, tstep, {"Cutoff=", threshold, "Maxm=", maxBondDim});
  // one problem per worker: every copy owns its slice store and its ControlBasis cache
  std::vector<ControlBasis> bases;
  for (int w = 0; w < workers; ++w) bases.push_back(ControlBasisFactory::buildChoppedSineBasis(u0, tstep, T, (size_t)M));
  std::vector<std::unique_ptr<OC_t>> problems;
  for (int w = 0; w < workers; ++w) problems.emplace_back(new OC_t(psi_f, psi_i, stepper, bases[w], gamma));
  std::vector<PenalisedCost> costs;
  for (int w = 0; w < workers; ++w) costs.push_back(PenalisedCost{*problems[w], 2.0, 100.0, gammaBound});

  Amoeba::Batch batch = [&](const std::vector<Amoeba::Point>& pts) {
    std::vector<double> out(pts.size());
    std::vector<std::thread> pool;
    for (int w = 0; w < workers; ++w)
      pool.emplace_back([&, w] { for (size_t k = (size_t)w; k < pts.size(); k += (size_t)workers) out[k] = costs[w](pts[k]); });
    for (auto& t : pool) t.join();
    return out;
  };

  std::valarray<double> x0((size_t)M);
  x0 = 0.0;
  Amoeba opt((size_t)M);
  opt.setLimits((unsigned)maxFun, 5000, std::max(optTol, 1e-12));
  auto result = opt.optimize(x0, costs[0], batch);
  const double best = std::get<0>(result);
  const std::valarray<double>& xb = std::get<1>(result);
  std::printf("\nBest cost %.12e\n", best);

  const std::vector<double> c0((size_t)M, 0.0), c1(std::begin(xb), std::end(xb));
  OC_t& OC = *problems[0];
  const auto f0 = OC.getFidelityForAllT(c0), f1 = OC.getFidelityForAllT(c1);
  const auto ui = OC.getControl(c0), uf = OC.getControl(c1);
  const auto times = OC.getTimeAxis();
  std::ofstream ramp("BHrampInitialFinal.txt");
  for (size_t i = 0; i < times.size(); ++i) ramp << times[i] << "\t" << ui[i] << "\t" << f0[i] << "\t" << uf[i] << "\t" << f1[i] << "\n";
  std::ofstream res("AmoebaResult.txt");
  res.precision(15);
  res << best << "\n";
  for (double v : c1) res << v << "\t";
  res << "\n";
  std::ofstream hist("AmoebaHistory.txt");
  const auto& ch = std::get<2>(result);
  const auto& eh = std::get<3>(result);
  for (size_t i = 0; i < ch.size(); ++i) {
    hist << i << "\t" << ch[i] << "\t" << eh[i] << "\n";
    if (i > 0 && eh[i] == eh[i - 1] && ch[i] == ch[i - 1] && i > 2) break;       // the tail only repeats the last entry
  }
  return 0;
}
