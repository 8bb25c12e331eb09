// Ramp optimisation driver on the GPU engine: what the reference's main/OptimizeRamp.cpp does (input file -> GROUP problem ->
// optimiser behind the TNLP adapter -> ExpectationN.txt), with the same input keys, defaults and output files.
//   usage: OptimizeRamp InputFile_BHcontrol [seed]
// Input keys (group "input", main/OptimizeRamp.cpp:27-50): tstep (1e-2), T, N, Npart, d, M, gamma (0), cacheProgress (no),
// useBFGS (no), maxBondDim (100), optTol (1e-7), threshold (1e-7), threadCount (2), maxIter (200), maxCPUHours (24), ObjScaling (1).
// Files written: BHrampInitialFinal.txt, GROUPHessian.txt, GRAPEHessian.txt (BH_nlp::finalize_solution), ProgressCache.txt
// (cacheProgress = yes), ExpectationN.txt (time, <N_1> .. <N_L> per time slice).
#include "BH_nlp.hpp"
#include "BH_tDMRG.hpp"
#include "ControlBasisFactory.hpp"
#include "InitializeState.hpp"
#include "IpIpoptApplication.hpp"
#include "OptimalControl.hpp"
#include "SeedGenerator.hpp"
#include "correlations.hpp"

#include <cstdio>
#include <cstdlib>
#include <fstream>
#include <iostream>
#include <string>

using namespace itensor;

struct RampInput {
  double tstep, T, gamma, optTol, threshold, maxCPUHours, ObjScaling;
  int N, Npart, d, M, maxBondDim, maxIter;
  size_t threadCount;
  bool cache, useBFGS;
  explicit RampInput(const InputGroup& in)
      : tstep(in.getReal("tstep", 1e-2)), T(in.getReal("T")), gamma(in.getReal("gamma", 0)), optTol(in.getReal("optTol", 1e-7)),
        threshold(in.getReal("threshold", 1e-7)), maxCPUHours(in.getReal("maxCPUHours", 24)), ObjScaling(in.getReal("ObjScaling", 1)),
        N(in.getInt("N")), Npart(in.getInt("Npart")), d(in.getInt("d")), M(in.getInt("M")), maxBondDim(in.getInt("maxBondDim", 100)),
        maxIter(in.getInt("maxIter", 200)), threadCount((size_t)in.getInt("threadCount", 2)), cache(in.getYesNo("cacheProgress", false)),
        useBFGS(in.getYesNo("useBFGS", false)) {}
  void print(int seed) const {
    std::printf("Optimal control of the Bose-Hubbard model on the GPU engine\n");
    std::printf("  sites %d, particles %d, local dimension d = %d\n  duration T = %g, time step %g, GROUP dimension M = %d, gamma = %g\n", N, Npart, d, T,
                tstep, M, gamma);
    std::printf("  max bond dimension %d, truncation threshold %g, BFGS %d\n  optimiser: tol %g, max iterations %d, max CPU time %g s, objective scaling %g\n",
                maxBondDim, threshold, (int)useBFGS, optTol, maxIter, maxCPUHours * 3600.0, ObjScaling);
    std::printf("  threadCount %zu, seed %d\n\n", threadCount, seed);
  }
};

static void write_expectation_n(const SiteSet& sites, OptimalControl<BH_tDMRG>& OC, const std::string& filename) {
  std::ofstream out(filename);
  if (!out.is_open()) { std::cout << "Unable to open file\n"; return; }
  auto slices = OC.getPsit();
  const auto times = OC.getTimeAxis();
  for (size_t i = 0; i < slices.size(); ++i) {
    out << times.at(i) << "\t";
    for (const auto& v : expectationValues(sites, slices[i], "N")) out << v.real() << "\t";
    out << "\n";
  }
}

int main(int argc, char* argv[]) {
  if (argc < 2) { std::printf("Usage: %s InputFile_BHcontrol [seed]\n", argv[0]); return 0; }
  const RampInput in{InputGroup(argv[1], "input")};
  const int seed = argc > 2 ? std::stoi(argv[2]) : 1;
  if (argc <= 2) std::printf("Default seed used\n");
  std::srand(123456789u * (unsigned)seed);                       // main/OptimizeRamp.cpp:60
  in.print(seed);
  const double J = 1.0, U_i = 2.5, U_f = 50;

  auto sites = BoseHubbard(in.N, in.d);
  auto u0 = SeedGenerator::linsigmoidSeed(U_i, U_f, (size_t)(in.T / in.tstep + 1));
  auto basis = ControlBasisFactory::buildChoppedSineBasis(u0, in.tstep, in.T, (size_t)in.M);
  auto psi_i = InitializeState(sites, in.Npart, J, u0.front(), in.maxBondDim, in.threshold);
  auto psi_f = InitializeState(sites, in.Npart, J, u0.back(), in.maxBondDim, in.threshold);
  auto stepper = BH_tDMRG(sites, J, in.tstep, {"Cutoff=", in.threshold, "Maxm=", in.maxBondDim});
  OptimalControl<BH_tDMRG> OC(psi_f, psi_i, stepper, basis, in.gamma, in.useBFGS);
  OC.setThreadCount(in.threadCount);

  Ipopt::SmartPtr<Ipopt::TNLP> nlp(new BH_nlp(OC, in.cache));
  auto app = Ipopt::IpoptApplicationFactory();
  app->Options()->SetNumericValue("tol", in.optTol);
  app->Options()->SetStringValue("mu_strategy", "adaptive");
  app->Options()->SetStringValue("jac_d_constant", "yes");
  app->Options()->SetIntegerValue("max_iter", in.maxIter);
  app->Options()->SetNumericValue("max_cpu_time", in.maxCPUHours * 3600.0);
  app->Options()->SetNumericValue("obj_scaling_factor", in.ObjScaling);
  if (in.useBFGS) app->Options()->SetStringValue("hessian_approximation", "limited-memory");
  if (app->Initialize() != Ipopt::Solve_Succeeded) { std::printf("\n\n*** Error during initialization!\n"); return 0; }
  const auto status = app->OptimizeTNLP(nlp);
  std::printf(status == Ipopt::Solve_Succeeded ? "\n\n*** The problem solved!\n" : "\n\n*** The problem FAILED!\n");
  std::printf("iterations %d, final objective %.12e\n", app->IterationCount(), app->FinalObjective());

  OC.setGRAPE(false);                                            // finalize_solution leaves the problem in GRAPE mode
  write_expectation_n(sites, OC, "ExpectationN.txt");
  return 0;
}
