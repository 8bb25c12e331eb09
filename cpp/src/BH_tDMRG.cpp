// BH_tDMRG over libocmps: the gates, the sweep order and the truncation live in the library
// (optimalcontrolmps_b200/csrc/engine.cu follows the reference's src/BH_tDMRG.cpp:18-230).
#include "BH_tDMRG.hpp"

#include <algorithm>
#include <cmath>
#include <map>
#include <mutex>

namespace {
// exactApplyMPO(K, psi, args) carries no stepper in its signature; steppers register by (Cutoff, Maxm) so that the
// call site src/OptimalControl.cpp:256 -- exactApplyMPO(stepper.propagatorDeriv(u), psi, stepper.getArgs()) -- resolves.
std::mutex g_reg_mutex;
std::vector<std::pair<std::pair<double, int>, ocmps_stepper*>> g_registry;
}  // namespace

BH_tDMRG::BH_tDMRG(const SiteSet& sites, const double J, const double tstep, const Args& args, int chi_cap)
    : p_(std::make_shared<Holder>()), sites_(sites), args_(args), J_(J), tstep_(tstep) {
  const int L = sites.N(), D = sites.D();
  const double cutoff = args.defined("Cutoff") ? args.getReal("Cutoff") : -1.0;
  const int maxm = args.defined("Maxm") ? args.getInt("Maxm") : 0;
  long long full = 1;
  for (int i = 0; i < L / 2 && full < 4096; ++i) full *= D;
  if (chi_cap <= 0) chi_cap = maxm > 0 ? maxm : (int)std::min<long long>(full, 256);
  cap_ = (int)std::min<long long>(chi_cap, full);
  ocmps_check(ocmps_stepper_create(default_context(), L, D, J, tstep, cutoff, maxm, cap_, 0, &p_->h), "ocmps_stepper_create");
  std::lock_guard<std::mutex> lock(g_reg_mutex);
  g_registry.push_back({{cutoff, maxm}, p_->h});
}

void BH_tDMRG::setTstep(const double t) {
  tstep_ = t;
  ocmps_check(ocmps_stepper_set_tstep(p_->h, t), "ocmps_stepper_set_tstep");
}

double BH_tDMRG::getTstep() const { return tstep_; }

Args BH_tDMRG::getArgs() const { return args_; }

IQMPO BH_tDMRG::propagatorDeriv(const double&) const { return IQMPO(IQMPO::PropagatorDerivative); }   // constant, argument unused

void BH_tDMRG::step(IQMPS& psi, const double from, const double to, bool propagateForward) const {
  if (psi.capacity() != cap_) psi = psi.withCapacity(cap_);
  ocmps_check(ocmps_step(p_->h, psi.handle(), from, to, propagateForward ? 1 : 0), "ocmps_step");
}

namespace itensor {
IQMPS exactApplyMPO(const IQMPO& K, const IQMPS& psi, const Args& args) {
  if (K.kind != IQMPO::PropagatorDerivative) throw std::invalid_argument("exactApplyMPO: only the propagator derivative MPO is supported");
  const double cutoff = args.defined("Cutoff") ? args.getReal("Cutoff") : -1.0;
  const int maxm = args.defined("Maxm") ? args.getInt("Maxm") : 0;
  ocmps_stepper* st = nullptr;
  {
    std::lock_guard<std::mutex> lock(g_reg_mutex);
    for (auto it = g_registry.rbegin(); it != g_registry.rend(); ++it)
      if (it->first.first == cutoff && it->first.second == maxm) { st = it->second; break; }
  }
  if (!st) throw std::invalid_argument("exactApplyMPO: no BH_tDMRG with these Args exists");
  IQMPS out(psi.N(), psi.D(), psi.capacity());
  ocmps_check(ocmps_apply_K(st, psi.handle(), out.handle()), "ocmps_apply_K");
  return out;
}
}  // namespace itensor
