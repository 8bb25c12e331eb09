// Control seeds (behavioural mirror of the reference's include/SeedGenerator.hpp; std only).
#ifndef OCMPS_SEEDGENERATOR_HPP
#define OCMPS_SEEDGENERATOR_HPP
#include <cmath>
#include <cstdlib>
#include <vector>

class SeedGenerator {
  static double uniform(double lo, double hi) { return lo + (double)rand() / RAND_MAX * (hi - lo); }
 public:
  // n points from a to b by repeated addition, with the reference's 1e-7 end guard (:26-37)
  static std::vector<double> linspace(double a, double b, int n) {
    std::vector<double> out;
    const double h = (b - a) / (n - 1);
    for (double x = a; x <= b + 1e-7; x += h) out.push_back(x);
    return out;
  }
  static std::vector<double> generateRange(double a, double step, double c) {
    std::vector<double> out;
    for (double x = a; x <= c + 1e-7; x += step) out.push_back(x);
    return out;
  }
  static std::vector<double> sigmoid(std::vector<double>& x, double k, double offset) {
    std::vector<double> s(x.size());
    for (size_t i = 0; i < x.size(); ++i) s[i] = 1.0 / (1.0 + std::exp(-k * (x[i] - offset)));
    return s;
  }
  // random linear+sigmoid ramp with pinned end behaviour (:66-95); coefficients from libc rand() like the reference
  static std::vector<double> linsigmoidSeed(double u_start, double u_end, size_t length) {
    std::vector<double> x = linspace(0, 100, (int)length);
    const double a = uniform(0.01, 0.15);
    const double b = u_end - u_start - a * x.back();
    const double c = uniform(0.06, 0.18);
    const double d = uniform(60, 80);
    std::vector<double> w = sigmoid(x, 0.7, 5), w2 = sigmoid(x, -0.9, 100 - 7);
    for (size_t i = w.size() / 2; i < w.size(); ++i) w[i] = w2[i];
    w.front() = 0;
    w.back() = 0;
    std::vector<double> u(x.size());
    for (size_t i = 0; i < x.size(); ++i) {
      const double inner = a * x[i] + b / (1 + std::exp(-c * (x[i] - d))) + u_start;
      const double outer = (u_end - u_start) / (1 + std::exp(-0.2 * (x[i] - 40))) + u_start;
      u[i] = w[i] * inner + (1 - w[i]) * outer;
    }
    return u;
  }
  static std::vector<double> adiabaticSeed(double u_start, double u_end, size_t length) {
    std::vector<double> u = linspace(0, 100, (int)length);
    const double p = 3.5, k = 1.0 / 3.0, xs = 40, a = 0.01;
    for (double& x : u)
      x = x < xs ? (p - u_start - a * xs) / (1 + std::exp(-k * (x - xs / 2.0))) + u_start + a * x
                 : std::exp(std::log(u_end - p + 1) / (100 - xs) * (x - xs)) + p - 1;
    return u;
  }
  static std::vector<double> randomCoeffSeed(double lo, double hi, size_t N) {
    std::vector<double> v(N);
    for (double& x : v) x = uniform(lo, hi);
    return v;
  }
};
#endif
