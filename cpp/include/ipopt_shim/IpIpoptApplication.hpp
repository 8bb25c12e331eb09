// Stand-in for IPOPT's application object (IpIpoptApplication.hpp), which main/OptimizeRamp.cpp:97-129 drives: options, Initialize(),
// OptimizeTNLP().  IPOPT is not available here, so OptimizeTNLP runs a small optimiser of its own behind the same calls, good enough
// to run the drivers end to end:
//   * box bounds on x by projection, the constraints g_l <= g(x) <= g_u (here: 2 <= u_i <= 100, linear in x) by backtracking;
//   * search direction: Newton with the exact Hessian from eval_h (Levenberg shift until positive definite), or L-BFGS when the
//     option hessian_approximation = limited-memory is set (the reference's BFGS mode, main/OptimizeRamp.cpp:109-111);
//   * Armijo backtracking; stops at the projected-gradient tolerance `tol`, `max_iter` or `max_cpu_time`;
//   * the TNLP callbacks are issued in IPOPT's order with IPOPT's new_x convention (first call at a new point has new_x = true,
//     src/BH_nlp.cpp:118-189), intermediate_callback once per iteration, finalize_solution at the end.
// With a real IPOPT installation put its include directory before this one.
#ifndef OCMPS_IPOPT_SHIM_APPLICATION_HPP
#define OCMPS_IPOPT_SHIM_APPLICATION_HPP
#include <algorithm>
#include <chrono>
#include <cmath>
#include <cstdio>
#include <deque>
#include <map>
#include <memory>
#include <string>
#include <vector>
#include "IpTNLP.hpp"

namespace Ipopt {

template <class T> using SmartPtr = std::shared_ptr<T>;       // `SmartPtr<TNLP> p = new BH_nlp(...)` is written `SmartPtr<TNLP> p(new ...)`

enum ApplicationReturnStatus { Solve_Succeeded = 0, Solved_To_Acceptable_Level = 1, Maximum_Iterations_Exceeded = -1,
                               Maximum_CpuTime_Exceeded = -4, Search_Direction_Becomes_Too_Small = -3, Internal_Error = -199 };

class OptionsList {
  std::map<std::string, double> num_;
  std::map<std::string, int> int_;
  std::map<std::string, std::string> str_;
 public:
  bool SetNumericValue(const std::string& k, double v) { num_[k] = v; return true; }
  bool SetIntegerValue(const std::string& k, int v) { int_[k] = v; return true; }
  bool SetStringValue(const std::string& k, const std::string& v) { str_[k] = v; return true; }
  double num(const std::string& k, double d) const { auto it = num_.find(k); return it == num_.end() ? d : it->second; }
  int integer(const std::string& k, int d) const { auto it = int_.find(k); return it == int_.end() ? d : it->second; }
  std::string str(const std::string& k, const std::string& d) const { auto it = str_.find(k); return it == str_.end() ? d : it->second; }
};

class IpoptApplication {
  std::shared_ptr<OptionsList> opts_ = std::make_shared<OptionsList>();
  int iterations_ = 0;
  double final_obj_ = 0.0;
  static double dot(const std::vector<double>& a, const std::vector<double>& b) { double s = 0; for (size_t i = 0; i < a.size(); ++i) s += a[i] * b[i]; return s; }
  // solves (H + shift I) d = -g by Cholesky; false if not positive definite
  static bool newton_dir(const std::vector<double>& Hlow, const std::vector<double>& g, double shift, std::vector<double>& d) {
    const int n = (int)g.size();
    std::vector<double> L((size_t)n * n, 0.0);
    for (int r = 0, k = 0; r < n; ++r) for (int c = 0; c <= r; ++c, ++k) L[(size_t)r * n + c] = Hlow[k];
    for (int j = 0; j < n; ++j) {
      double s = L[(size_t)j * n + j] + shift;
      for (int k = 0; k < j; ++k) s -= L[(size_t)j * n + k] * L[(size_t)j * n + k];
      if (!(s > 1e-14)) return false;
      const double ljj = std::sqrt(s);
      L[(size_t)j * n + j] = ljj;
      for (int i = j + 1; i < n; ++i) {
        double t = L[(size_t)i * n + j];
        for (int k = 0; k < j; ++k) t -= L[(size_t)i * n + k] * L[(size_t)j * n + k];
        L[(size_t)i * n + j] = t / ljj;
      }
    }
    std::vector<double> y(n);
    for (int i = 0; i < n; ++i) { double t = -g[i]; for (int k = 0; k < i; ++k) t -= L[(size_t)i * n + k] * y[k]; y[i] = t / L[(size_t)i * n + i]; }
    d.assign(n, 0.0);
    for (int i = n - 1; i >= 0; --i) { double t = y[i]; for (int k = i + 1; k < n; ++k) t -= L[(size_t)k * n + i] * d[k]; d[i] = t / L[(size_t)i * n + i]; }
    return true;
  }
 public:
  std::shared_ptr<OptionsList> Options() { return opts_; }
  ApplicationReturnStatus Initialize() { return Solve_Succeeded; }
  int IterationCount() const { return iterations_; }
  double FinalObjective() const { return final_obj_; }

  ApplicationReturnStatus OptimizeTNLP(const SmartPtr<TNLP>& nlp) {
    Index n = 0, m = 0, nnzj = 0, nnzh = 0;
    TNLP::IndexStyleEnum style;
    if (!nlp->get_nlp_info(n, m, nnzj, nnzh, style)) return Internal_Error;
    std::vector<double> xl(n), xu(n), gl(m), gu(m), x(n), g(m), grad(n), zl(n, 0.0), zu(n, 0.0), lam(m, 0.0);
    nlp->get_bounds_info(n, xl.data(), xu.data(), m, gl.data(), gu.data());
    nlp->get_starting_point(n, true, x.data(), false, nullptr, nullptr, m, false, nullptr);
    const double tol = opts_->num("tol", 1e-8), scale = opts_->num("obj_scaling_factor", 1.0), max_cpu = opts_->num("max_cpu_time", 1e20);
    const int max_iter = opts_->integer("max_iter", 3000);
    const bool lbfgs = opts_->str("hessian_approximation", "exact") == "limited-memory";
    const auto t0 = std::chrono::steady_clock::now();
    auto elapsed = [&] { return std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count(); };
    auto feasible = [&](const std::vector<double>& gv) {
      for (Index i = 0; i < m; ++i) if (gv[i] < gl[i] - 1e-12 || gv[i] > gu[i] + 1e-12) return false;
      return true;
    };
    double f = 0.0;
    nlp->eval_g(n, x.data(), true, m, g.data());                 // (propagates psi for x, like IPOPT's first constraint evaluation)
    nlp->eval_f(n, x.data(), false, f);
    nlp->eval_grad_f(n, x.data(), false, grad.data());
    std::deque<std::pair<std::vector<double>, std::vector<double>>> mem;     // L-BFGS pairs (s, y)
    std::vector<double> hvals(nnzh), d(n), xn(n), gn(m), gradn(n);
    ApplicationReturnStatus status = Maximum_Iterations_Exceeded;
    int ls_trials = 0;
    for (iterations_ = 0;; ++iterations_) {
      double pg = 0.0;                                           // projected gradient (box constraints on x)
      for (Index i = 0; i < n; ++i) {
        double gi = grad[i];
        if ((x[i] <= xl[i] && gi > 0) || (x[i] >= xu[i] && gi < 0)) gi = 0;
        pg = std::max(pg, std::fabs(gi) * scale);
      }
      if (!nlp->intermediate_callback(RegularMode, iterations_, f, 0.0, pg, 0.0, 0.0, 0.0, 1.0, 1.0, ls_trials, nullptr, nullptr)) { status = Internal_Error; break; }
      std::printf("iter %3d  f = %.10e  |proj grad| = %.3e  ls %d\n", iterations_, f, pg, ls_trials);
      if (pg <= tol) { status = Solve_Succeeded; break; }
      if (iterations_ >= max_iter) { status = Maximum_Iterations_Exceeded; break; }
      if (elapsed() > max_cpu) { status = Maximum_CpuTime_Exceeded; break; }
      // ---- direction ----
      bool have = false;
      if (!lbfgs && nnzh > 0) {
        std::vector<Index> ir(nnzh), jc(nnzh);
        if (nlp->eval_h(n, x.data(), false, 1.0, m, lam.data(), false, nnzh, nullptr, nullptr, hvals.data())) {
          double shift = 0.0;
          for (int tries = 0; tries < 40 && !have; ++tries) {
            have = newton_dir(hvals, grad, shift, d);
            if (!have || dot(d, grad) >= 0) { have = false; shift = shift == 0.0 ? 1e-8 : shift * 10.0; }
          }
        }
      } else if (!mem.empty()) {
        std::vector<double> q = grad, alpha(mem.size());
        for (int i = (int)mem.size() - 1; i >= 0; --i) {
          alpha[i] = dot(mem[i].first, q) / dot(mem[i].second, mem[i].first);
          for (Index k = 0; k < n; ++k) q[k] -= alpha[i] * mem[i].second[k];
        }
        const double gam = dot(mem.back().first, mem.back().second) / dot(mem.back().second, mem.back().second);
        for (Index k = 0; k < n; ++k) q[k] *= gam;
        for (size_t i = 0; i < mem.size(); ++i) {
          const double beta = dot(mem[i].second, q) / dot(mem[i].second, mem[i].first);
          for (Index k = 0; k < n; ++k) q[k] += (alpha[i] - beta) * mem[i].first[k];
        }
        for (Index k = 0; k < n; ++k) d[k] = -q[k];
        have = dot(d, grad) < 0;
      }
      if (!have) { const double gnorm = std::sqrt(dot(grad, grad)); for (Index k = 0; k < n; ++k) d[k] = -grad[k] / std::max(gnorm, 1e-300); }
      // ---- backtracking: box by projection, g-constraints and Armijo by shrinking ----
      double alpha = 1.0, fn = f;
      bool accepted = false;
      ls_trials = 0;
      for (; ls_trials < 30; ++ls_trials, alpha *= 0.5) {
        for (Index k = 0; k < n; ++k) xn[k] = std::min(xu[k], std::max(xl[k], x[k] + alpha * d[k]));
        nlp->eval_g(n, xn.data(), true, m, gn.data());
        if (!feasible(gn)) continue;
        nlp->eval_f(n, xn.data(), false, fn);
        double slope = 0.0;
        for (Index k = 0; k < n; ++k) slope += grad[k] * (xn[k] - x[k]);
        if (fn <= f + 1e-4 * slope) { accepted = true; break; }
      }
      if (!accepted) { status = Search_Direction_Becomes_Too_Small; break; }
      nlp->eval_grad_f(n, xn.data(), false, gradn.data());
      std::vector<double> s(n), y(n);
      for (Index k = 0; k < n; ++k) { s[k] = xn[k] - x[k]; y[k] = gradn[k] - grad[k]; }
      if (dot(s, y) > 1e-14 * std::sqrt(dot(s, s) * dot(y, y))) { mem.push_back({s, y}); if (mem.size() > 8) mem.pop_front(); }
      x = xn; g = gn; grad = gradn; f = fn;
    }
    final_obj_ = f;
    nlp->finalize_solution(status == Solve_Succeeded ? SUCCESS : MAXITER_EXCEEDED, n, x.data(), zl.data(), zu.data(), m, g.data(), lam.data(), f, nullptr,
                           nullptr);
    return status;
  }
};

inline SmartPtr<IpoptApplication> IpoptApplicationFactory() { return std::make_shared<IpoptApplication>(); }

}  // namespace Ipopt
#endif
