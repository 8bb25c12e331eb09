// Observables of the reference's include/correlations.hpp that are diagonal in the boson number, on the device engine:
//   expectationValue(sites, psi, opname, i)  (:99-107)   and   expectationValues(sites, psi, opname)  (:109-117)
// with opname "N", "N(N-1)" or "NN" (include/BH_sites.h:129-171).  The value is <psi|O_i|psi>, not divided by the norm,
// as in the reference.  The state is copied into a one-slot slice store and evaluated by the batched transfer-matrix
// chain (ocmps_store_site_expectations); that chain needs the orthogonality centre at site 1, which every state that
// went through BH_tDMRG::step has -- any other gauge is detected and reported.
// entanglementEntropy(sites, psi) (:119-148) is provided as well (gauge moves of the engine, spectrum per bond).
// Not provided (ITensor-only post-processing): correlationFunction/Matrix/Term.
#ifndef OCMPS_CORRELATIONS_HPP
#define OCMPS_CORRELATIONS_HPP

#include <cmath>
#include <stdexcept>
#include <string>
#include <vector>

#include "itensor/all.h"

namespace ocmps_detail {
inline std::vector<double> site_op_diagonal(const std::string& opname, int D) {
  std::vector<double> v(D);
  for (int n = 0; n < D; ++n) {
    if (opname == "N") v[n] = n;
    else if (opname == "N(N-1)") v[n] = n * (n - 1.0);
    else if (opname == "NN") v[n] = (double)n * n;
    else if (opname == "Id") v[n] = 1.0;
    else throw std::invalid_argument("correlations.hpp: operator '" + opname + "' is not diagonal in the boson number");
  }
  return v;
}
}  // namespace ocmps_detail

inline std::vector<itensor::Cplx> expectationValues(itensor::SiteSet const& sites, itensor::IQMPS& psi, std::string const& opname) {
  using namespace itensor;
  const int L = psi.N(), D = psi.D();
  (void)sites;
  ocmps_store* store = nullptr;
  ocmps_check(ocmps_store_create(default_context(), L, D, psi.capacity(), 1, &store), "ocmps_store_create");
  std::vector<double> out(L), nrm(L);
  const std::vector<double> diag = ocmps_detail::site_op_diagonal(opname, D);
  int rc = ocmps_store_put(store, 0, psi.handle());
  if (!rc) rc = ocmps_store_site_expectations(store, 0, 1, diag.data(), 1, out.data(), nrm.data());
  ocmps_store_destroy(store);
  ocmps_check(rc, "ocmps_store_site_expectations");
  for (int j = 1; j < L; ++j)
    if (std::fabs(nrm[j] - nrm[0]) > 1e-9 * std::fabs(nrm[0]))
      throw std::runtime_error("expectationValues: the orthogonality centre of psi is not at site 1");
  std::vector<Cplx> res(L);
  for (int j = 0; j < L; ++j) res[j] = Cplx(out[j], 0.0);
  return res;
}

inline itensor::Cplx expectationValue(itensor::SiteSet const& sites, itensor::IQMPS& psi, std::string const& opname, int i) {
  return expectationValues(sites, psi, opname).at(i - 1);       // sites are 1-based in the reference
}

inline std::vector<double> entanglementEntropy(itensor::SiteSet const& sites, itensor::IQMPS& psi) {
  using namespace itensor;
  const int L = psi.N(), D = psi.D();
  (void)sites;
  ocmps_store* store = nullptr;
  ocmps_check(ocmps_store_create(default_context(), L, D, psi.capacity(), 1, &store), "ocmps_store_create");
  std::vector<double> S(L > 1 ? L - 1 : 0);
  int rc = ocmps_store_put(store, 0, psi.handle());
  if (!rc && L > 1) rc = ocmps_store_entanglement_entropy(store, 0, 1, S.data());
  ocmps_store_destroy(store);
  ocmps_check(rc, "ocmps_store_entanglement_entropy");
  return S;
}

#endif
