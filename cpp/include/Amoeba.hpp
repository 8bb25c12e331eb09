// Derivative-free Nelder-Mead search ("Amoeba"), behaviour of the reference's include/Amoeba.hpp:25-217: the same simplex
// construction (5 % perturbation, 0.00025 for zero entries :68-84), coefficients (rho 1, chi 2, psi 0.5, sigma 0.5 :36-41), move
// logic (:152-200), stopping rules (5000 evaluations / iterations or simplex cost spread <= 1e-6, :93-106) and return value
// (best cost, best point, cost history, evaluation-count history).  Addition for the GPU engine: the n evaluations that build
// the simplex and the n evaluations of a shrink are independent, so optimize() accepts an optional batch evaluator that gets
// all of them at once (the multi-seed / batched-controls path of the north star); the single-point callable is used elsewhere.
#pragma once
#include <algorithm>
#include <cmath>
#include <functional>
#include <iostream>
#include <stdexcept>
#include <string>
#include <tuple>
#include <valarray>
#include <vector>

struct Member {
  std::valarray<double> x;
  double fx;
  Member(std::valarray<double> x_, double fx_) : x(std::move(x_)), fx(fx_) {}
  friend bool operator<(const Member& a, const Member& b) { return a.fx < b.fx; }
  friend bool operator>(const Member& a, const Member& b) { return a.fx > b.fx; }
};

class Amoeba {
 public:
  using Point = std::valarray<double>;
  using Batch = std::function<std::vector<double>(const std::vector<Point>&)>;
  using Result = std::tuple<double, Point, std::valarray<double>, std::valarray<unsigned int>>;

  explicit Amoeba(std::size_t dimension) : dim_(dimension) { simplex_.reserve(dimension + 1); }
  void setDisplay(bool on) { display_ = on; }
  void setLimits(unsigned int maxFun, unsigned int maxIter, double tolFun) { maxFun_ = maxFun; maxIter_ = maxIter; tolFun_ = tolFun; }

  template <typename F> Result optimize(Point x0, F& f) { return optimize(x0, f, Batch()); }

  template <typename F> Result optimize(Point x0, F& f, const Batch& batch) {
    if (x0.size() != dim_) throw std::invalid_argument("x0 does not have the correct size.");
    unsigned int evals = 0, iter = 0;
    std::valarray<double> costs(maxIter_ + 1);
    std::valarray<unsigned int> counts(maxIter_ + 1);
    auto one = [&](const Point& p) { ++evals; return f(p); };
    auto many = [&](const std::vector<Point>& ps) {
      std::vector<double> out;
      if (batch) { out = batch(ps); evals += (unsigned int)ps.size(); }
      else for (const Point& p : ps) out.push_back(one(p));
      return out;
    };
    simplex_.clear();
    const double f0 = one(x0);
    if (display_) { std::cout << "Iteration\tFunc_evals\tBest\t\t  Action\n"; report(iter, evals, f0, "Start"); }
    costs[iter] = f0;
    {   // initial simplex: one vertex per coordinate
      std::vector<Point> ps;
      for (std::size_t i = 0; i < dim_; ++i) {
        Point p = x0;
        p[i] = p[i] != 0 ? (1 + usual_delta_) * p[i] : zero_term_delta_;
        ps.push_back(p);
      }
      const std::vector<double> fs = many(ps);
      simplex_.emplace_back(x0, f0);
      for (std::size_t i = 0; i < dim_; ++i) simplex_.emplace_back(ps[i], fs[i]);
    }
    ++iter;
    std::sort(simplex_.begin(), simplex_.end());
    if (display_) report(iter, evals, simplex_[0].fx, "Initialize");
    costs[iter] = simplex_[0].fx;
    counts[iter] = evals;
    while (!stop(evals, iter)) {
      Point bar = simplex_[0].x;
      for (std::size_t i = 1; i < dim_; ++i) bar = bar + simplex_[i].x;
      bar = (1.0 / (double)dim_) * bar;
      Member& worst = simplex_[dim_];
      const Point xr = (1.0 + rho_) * bar - rho_ * worst.x;
      const double fr = one(xr);
      std::string what;
      bool shrink = false;
      if (fr < simplex_[0].fx) {
        const Point xe = (1 + rho_ * chi_) * bar - rho_ * chi_ * worst.x;
        const double fe = one(xe);
        if (fe < fr) { worst = Member(xe, fe); what = "Expand"; } else { worst = Member(xr, fr); what = "Reflect"; }
      } else if (fr < simplex_[dim_ - 1].fx) {
        worst = Member(xr, fr);
        what = "Reflect";
      } else if (fr < worst.fx) {
        const Point xc = (1 + psi_ * rho_) * bar - psi_ * rho_ * worst.x;
        const double fc = one(xc);
        if (fc <= fr) { worst = Member(xc, fc); what = "Contract outside"; } else shrink = true;
      } else {
        const Point xcc = (1 - psi_) * bar + psi_ * worst.x;
        const double fcc = one(xcc);
        if (fcc < worst.fx) { worst = Member(xcc, fcc); what = "Contract inside"; } else shrink = true;
      }
      if (shrink) {        // towards vertex 1, as the reference does (:86-91)
        std::vector<Point> ps;
        for (std::size_t i = 1; i <= dim_; ++i) ps.push_back(simplex_[1].x + sigma_ * (simplex_[i].x - simplex_[1].x));
        const std::vector<double> fs = many(ps);
        for (std::size_t i = 1; i <= dim_; ++i) simplex_[i] = Member(ps[i - 1], fs[i - 1]);
        what = "Shrink";
      }
      std::sort(simplex_.begin(), simplex_.end());
      ++iter;
      if (display_) report(iter, evals, simplex_[0].fx, what);
      costs[iter] = simplex_[0].fx;
      counts[iter] = evals;
    }
    for (std::size_t i = iter; i <= maxIter_; ++i) { costs[i] = simplex_[0].fx; counts[i] = evals; }
    return std::make_tuple(simplex_[0].fx, simplex_[0].x, costs, counts);
  }

 private:
  std::size_t dim_;
  std::vector<Member> simplex_;
  bool display_ = true;
  unsigned int maxFun_ = 5000, maxIter_ = 5000;
  double tolFun_ = 1e-6;
  double usual_delta_ = 0.05, zero_term_delta_ = 0.00025, rho_ = 1.0, chi_ = 2.0, psi_ = 0.5, sigma_ = 0.5;

  bool stop(unsigned int evals, unsigned int iter) const {
    if (evals >= maxFun_ || iter >= maxIter_) return true;
    double spread = 0.0;
    for (std::size_t i = 0; i < dim_; ++i) spread = std::max(spread, std::abs(simplex_[0].fx - simplex_[i + 1].fx));
    return spread <= tolFun_;
  }
  static void report(unsigned int iter, unsigned int evals, double best, const std::string& what) {
    std::cout << "    " << iter << "\t\t    " << evals << "\t      " << best << "\t\t" << what << std::endl;
  }
};
