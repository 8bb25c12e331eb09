// BH_tDMRG over libocmps: the gates, the sweep order and the truncation live in the library
// (optimalcontrolmps_b200/csrc/engine.cu follows the reference's src/BH_tDMRG.cpp:18-230).
#include "BH_tDMRG.hpp"

#include <algorithm>
#include <cmath>
#include <map>
#include <mutex>

namespace {
// exactApplyMPO(K, psi, args) carries no stepper in its signature.  The MPO returned by propagatorDeriv names the stepper
// that built it (the reference's call site is exactApplyMPO(stepper.propagatorDeriv(u), psi, stepper.getArgs()),
// src/OptimalControl.cpp:256); an MPO without one falls back to the live steppers registered here, matched on the full
// shape (L, D, capacity) and Args.  Entries are weak: a destroyed stepper can never be handed out.
struct RegEntry { int L, D, cap; double cutoff; int maxm; std::weak_ptr<void> owner; ocmps_stepper* h; };
std::mutex g_reg_mutex;
std::vector<RegEntry> g_registry;
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
  g_registry.erase(std::remove_if(g_registry.begin(), g_registry.end(), [](const RegEntry& e) { return e.owner.expired(); }),
                   g_registry.end());
  g_registry.push_back({L, D, cap_, cutoff, maxm, std::weak_ptr<void>(p_), p_->h});
}

void BH_tDMRG::setTstep(const double t) {
  tstep_ = t;
  ocmps_check(ocmps_stepper_set_tstep(p_->h, t), "ocmps_stepper_set_tstep");
}

double BH_tDMRG::getTstep() const { return tstep_; }

Args BH_tDMRG::getArgs() const { return args_; }

IQMPO BH_tDMRG::propagatorDeriv(const double&) const {      // constant, argument unused (src/BH_tDMRG.cpp:238-241)
  IQMPO K(IQMPO::PropagatorDerivative);
  K.owner = p_;
  K.stepper = p_ ? p_->h : nullptr;
  K.cutoff = args_.defined("Cutoff") ? args_.getReal("Cutoff") : -1.0;
  K.maxm = args_.defined("Maxm") ? args_.getInt("Maxm") : 0;
  return K;
}

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
  std::shared_ptr<void> keep;                   // holds the stepper alive for the duration of the call
  if (K.stepper && K.owner && K.cutoff == cutoff && K.maxm == maxm) {
    st = K.stepper;
    keep = K.owner;
  } else {
    std::lock_guard<std::mutex> lock(g_reg_mutex);
    for (auto it = g_registry.rbegin(); it != g_registry.rend(); ++it) {
      if (it->L != psi.N() || it->D != psi.D() || it->cap != psi.capacity() || it->cutoff != cutoff || it->maxm != maxm) continue;
      keep = it->owner.lock();
      if (keep) { st = it->h; break; }
    }
  }
  if (!st) throw std::invalid_argument("exactApplyMPO: no live BH_tDMRG with this shape (L, D, capacity) and these Args exists");
  IQMPS out(psi.N(), psi.D(), psi.capacity());
  ocmps_check(ocmps_apply_K(st, psi.handle(), out.handle()), "ocmps_apply_K");
  return out;
}
}  // namespace itensor
