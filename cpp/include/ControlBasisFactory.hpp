// Chopped-sine GROUP basis (reference include/ControlBasisFactory.hpp:25-52, including its truncated PI).
#ifndef OCMPS_CONTROLBASISFACTORY_HPP
#define OCMPS_CONTROLBASISFACTORY_HPP
#include <cmath>
#include "ControlBasis.hpp"
#include "SeedGenerator.hpp"

class ControlBasisFactory {
 public:
  static ControlBasis buildChoppedSineBasis(stdvec& u0, double tstep, double T, size_t M) {
    const double pi_ref = 3.14159265;                        // the reference's "#define PI 3.14159265" is behaviour
    const size_t N = u0.size();
    stdvec x = SeedGenerator::linspace(0, 100, (int)N);
    stdvec S = SeedGenerator::sigmoid(x, 8.0, 1.1), S2 = SeedGenerator::sigmoid(x, -8.0, 100 - 1.1);
    for (size_t i = N / 2; i < N; ++i) S[i] = S2[i];
    S[0] = 0;
    S[N - 1] = 0;
    rowmat f(N, stdvec(M, 0.0));
    for (size_t i = 0; i < N; ++i)
      for (size_t n = 0; n < M; ++n) f[i][n] = std::sin((n + 1) * pi_ref * tstep * i / T);
    return ControlBasis(u0, S, f);
  }
};
#endif
