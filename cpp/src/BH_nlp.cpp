// Ipopt::TNLP adapter (behaviour of the reference's src/BH_nlp.cpp): n = M coefficients in [-20, 20], m = N
// constraints 2 <= u_i <= 100, dense Jacobian = control Jacobian, lower-triangular dense Hessian; IPOPT's
// new_x is handed straight through as new_control, eval_h always recomputes.
#include "BH_nlp.hpp"

#include <cassert>
#include <fstream>
#include <iostream>

BH_nlp::BH_nlp(OC_BH& prob, bool cache) : optControlProb(prob), cacheProgress(cache) { times = optControlProb.getTimeAxis(); }
BH_nlp::~BH_nlp() {}

bool BH_nlp::get_nlp_info(Ipopt::Index& n, Ipopt::Index& m, Ipopt::Index& nnz_jac_g, Ipopt::Index& nnz_h_lag, IndexStyleEnum& style) {
  n = (Ipopt::Index)optControlProb.getM();
  m = (Ipopt::Index)optControlProb.getN();
  nnz_jac_g = m * n;
  nnz_h_lag = (n * n + n) / 2;
  style = TNLP::C_STYLE;
  return true;
}

bool BH_nlp::get_bounds_info(Ipopt::Index n, Number* x_l, Number* x_u, Ipopt::Index m, Number* g_l, Number* g_u) {
  for (Ipopt::Index i = 0; i < n; ++i) { x_l[i] = -20; x_u[i] = 20; }
  for (Ipopt::Index i = 0; i < m; ++i) { g_l[i] = 2.0; g_u[i] = 100; }
  return true;
}

bool BH_nlp::get_starting_point(Ipopt::Index n, bool init_x, Number* x, bool init_z, Number*, Number*, Ipopt::Index, bool init_lambda,
                                Number*) {
  assert(init_x && !init_z && !init_lambda);
  (void)init_x; (void)init_z; (void)init_lambda;
  for (Ipopt::Index i = 0; i < n; ++i) x[i] = 0;
  initialCoeffs.assign(x, x + n);
  return true;
}

bool BH_nlp::eval_f(Ipopt::Index n, const Number* x, bool new_x, Number& obj) {
  obj = optControlProb.getCost(std::vector<double>(x, x + n), new_x);
  return true;
}

bool BH_nlp::eval_grad_f(Ipopt::Index n, const Number* x, bool new_x, Number* grad_f) {
  const auto g = optControlProb.getAnalyticGradient(std::vector<double>(x, x + n), new_x);
  std::copy(g.begin(), g.end(), grad_f);
  return true;
}

bool BH_nlp::eval_g(Ipopt::Index n, const Number* x, bool new_x, Ipopt::Index, Number* g) {
  const std::vector<double> c(x, x + n);
  if (new_x) optControlProb.propagatePsi(c);          // psi_t must match x for the later new_x=false calls
  const auto u = optControlProb.getControl(c);
  std::copy(u.begin(), u.end(), g);
  return true;
}

bool BH_nlp::eval_jac_g(Ipopt::Index n, const Number* x, bool new_x, Ipopt::Index m, Ipopt::Index, Ipopt::Index* iRow,
                        Ipopt::Index* jCol, Number* values) {
  if (new_x) optControlProb.propagatePsi(std::vector<double>(x, x + n));
  if (values == NULL) {
    for (Ipopt::Index i = 0; i < m; ++i)
      for (Ipopt::Index j = 0; j < n; ++j) { iRow[n * i + j] = i; jCol[n * i + j] = j; }
  } else {
    size_t k = 0;
    for (const auto& row : optControlProb.getControlJacobian())
      for (double v : row) values[k++] = v;
  }
  return true;
}

bool BH_nlp::eval_h(Ipopt::Index n, const Number* x, bool, Number obj_factor, Ipopt::Index, const Number*, bool, Ipopt::Index nele_hess,
                    Ipopt::Index* iRow, Ipopt::Index* jCol, Number* values) {
  Ipopt::Index k = 0;
  if (values == NULL) {
    for (Ipopt::Index r = 0; r < n; ++r)
      for (Ipopt::Index c = 0; c <= r; ++c) { iRow[k] = r; jCol[k] = c; ++k; }
    assert(k == nele_hess);
    (void)nele_hess;
  } else {
    const auto H = optControlProb.getHessian(std::vector<double>(x, x + n));      // always new_control = true
    for (Ipopt::Index r = 0; r < n; ++r)
      for (Ipopt::Index c = 0; c <= r; ++c) values[k++] = obj_factor * H[r][c];
  }
  return true;
}

static void write_matrix(const std::string& name, const rowmat& A) {
  std::ofstream f(name);
  if (!f.is_open()) { std::cout << "Unable to open file\n"; return; }
  for (const auto& row : A) { for (double v : row) f << v << "\t"; f << "\n"; }
}

void BH_nlp::finalize_solution(SolverReturn, Ipopt::Index n, const Number* x, const Number* z_L, const Number* z_U, Ipopt::Index,
                               const Number*, const Number*, Number obj_value, const IpoptData*, IpoptCalculatedQuantities*) {
  printf("\n\nSolution of the primal variables, x\n");
  for (Ipopt::Index i = 0; i < n; ++i) printf("x[%d] = %e\n", i, x[i]);
  printf("\n\nSolution of the bound multipliers, z_L and z_U\n");
  for (Ipopt::Index i = 0; i < n; ++i) printf("z_L[%d] = %e\n", i, z_L[i]);
  for (Ipopt::Index i = 0; i < n; ++i) printf("z_U[%d] = %e\n", i, z_U[i]);
  printf("\n\nObjective value\nf(x*) = %e\n", obj_value);
  const std::vector<double> fin(x, x + n);
  const auto u0 = optControlProb.getControl(initialCoeffs), u1 = optControlProb.getControl(fin);
  const auto f0 = optControlProb.getFidelityForAllT(initialCoeffs), f1 = optControlProb.getFidelityForAllT(fin);
  std::ofstream ramp("BHrampInitialFinal.txt");           // five tab-separated columns per time point
  if (ramp.is_open()) {
    for (size_t i = 0; i < times.size(); ++i)
      ramp << times.at(i) << "\t" << u0.at(i) << "\t" << f0.at(i) << "\t" << u1.at(i) << "\t" << f1.at(i) << "\n";
  } else {
    std::cout << "Unable to open file\n";
  }
  // The reference computes the final Hessians unconditionally (src/BH_nlp.cpp:258-261), which in BFGS mode reads stores that were
  // never allocated (README.md:5).  Here the xi / K.xi stores are switched back on first, so the two files are always written.
  if (optControlProb.useBFGS()) optControlProb.setBFGS(false);
  const auto Hgroup = optControlProb.getHessian(fin);
  optControlProb.setGRAPE(true);
  const auto Hgrape = optControlProb.getHessian(u1);
  write_matrix("GROUPHessian.txt", Hgroup);
  write_matrix("GRAPEHessian.txt", Hgrape);
}

bool BH_nlp::intermediate_callback(AlgorithmMode, Ipopt::Index iter, Number obj_value, Number, Number, Number, Number, Number, Number,
                                   Number, Ipopt::Index ls_trials, const IpoptData*, IpoptCalculatedQuantities*) {
  if (cacheProgress) {
    std::ofstream out("ProgressCache.txt", std::ios_base::app);
    if (out.is_open()) {
      const std::size_t steps = optControlProb.getN();
      std::size_t nprop = steps * (2 + ls_trials);                     // propagation-count model of the reference
      if (!optControlProb.useBFGS()) nprop += steps * (steps - 1) / 2;
      out << iter << "\t" << obj_value << "\t" << times.back() << "\t" << nprop << "\n";
    } else {
      std::cout << "Unable to open file\n";
    }
  }
  return true;
}
