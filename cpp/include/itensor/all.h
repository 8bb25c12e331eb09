// Minimal "itensor/all.h" compatibility shim over the libocmps C ABI.
//
// The reference's hot-path classes are written against ITensor v2 (include/OptimalControl.hpp:4,
// include/BH_tDMRG.hpp:4).  This header provides just the names those signatures use -- IQMPS, IQMPO,
// SiteSet / BoseHubbard, Args, Cplx, overlap / overlapC / norm / exactApplyMPO, linkInd(psi,b).m() -- on top
// of device-resident handles, so that callers such as main/OptimizeRamp.cpp:82-90 or src/BH_nlp.cpp keep
// compiling unchanged.  Nothing here computes on the CPU.
#ifndef OCMPS_ITENSOR_COMPAT_ALL_H
#define OCMPS_ITENSOR_COMPAT_ALL_H

#include <complex>
#include <map>
#include <memory>
#include <stdexcept>
#include <string>
#include <vector>

#include "ocmps.h"

namespace itensor {

using Real = double;
using Cplx = std::complex<double>;
static const Cplx Cplx_i(0.0, 1.0);

inline void ocmps_check(int rc, const char* what) {
  if (rc != 0) throw std::runtime_error(std::string(what) + ": " + ocmps_last_error());
}

// one context per GPU, created on first use (device chosen by OCMPS_DEVICE or 0)
ocmps_ctx* default_context();

// ---- Args{"Cutoff=",1E-8,"Maxm=",100} ----
class Args {
  std::map<std::string, double> vals_;
  static std::string key(std::string k) { if (!k.empty() && k.back() == '=') k.pop_back(); return k; }
  void fill() {}
  template <typename V, typename... Rest> void fill(const char* k, V v, Rest... rest) { vals_[key(k)] = (double)v; fill(rest...); }
 public:
  Args() {}
  template <typename V, typename... Rest> Args(const char* k, V v, Rest... rest) { fill(k, v, rest...); }
  bool defined(const std::string& k) const { return vals_.count(key(k)) > 0; }
  double getReal(const std::string& k, double def = 0.0) const { auto it = vals_.find(key(k)); return it == vals_.end() ? def : it->second; }
  int getInt(const std::string& k, int def = 0) const { auto it = vals_.find(key(k)); return it == vals_.end() ? def : (int)it->second; }
  void add(const std::string& k, double v) { vals_[key(k)] = v; }
};

// ---- InputGroup(file, "input"): the `name = value` blocks the reference's main programs read (main/OptimizeRamp.cpp:27-50) ----
//   input
//   {
//       N = 5
//       useBFGS = no
//   }
class InputGroup {
  std::map<std::string, std::string> kv_;
  std::string file_, group_;
  const std::string* find(const std::string& k) const { auto it = kv_.find(k); return it == kv_.end() ? nullptr : &it->second; }
 public:
  InputGroup(const std::string& filename, const std::string& groupname);
  bool has(const std::string& k) const { return kv_.count(k) > 0; }
  double getReal(const std::string& k) const;                      // mandatory: throws std::runtime_error if absent
  double getReal(const std::string& k, double def) const { return has(k) ? getReal(k) : def; }
  int getInt(const std::string& k) const;
  int getInt(const std::string& k, int def) const { return has(k) ? getInt(k) : def; }
  bool getYesNo(const std::string& k) const;
  bool getYesNo(const std::string& k, bool def) const { return has(k) ? getYesNo(k) : def; }
  std::string getString(const std::string& k) const;
  std::string getString(const std::string& k, const std::string& def) const { return has(k) ? getString(k) : def; }
};

// ---- SiteSet / BoseHubbard(N, d): N sites with occupations 0..d (include/BH_sites.h:10-58) ----
class SiteSet {
 protected:
  int N_ = 0, d_ = 0;
 public:
  SiteSet() {}
  SiteSet(int N, int d) : N_(N), d_(d) {}
  int N() const { return N_; }
  int d() const { return d_; }          // maximum occupation
  int D() const { return d_ + 1; }      // local dimension
};
class BoseHubbard : public SiteSet {
 public:
  BoseHubbard() {}
  BoseHubbard(int N, int d) : SiteSet(N, d) {}
};

struct LinkDim { long m_; long m() const { return m_; } };

// ---- IQMPS: value-semantic handle of a device-resident, charge-labelled MPS ----
class IQMPS {
  struct Holder { ocmps_mps* h = nullptr; ~Holder() { if (h) ocmps_mps_destroy(h); } };
  std::shared_ptr<Holder> p_;
  int L_ = 0, D_ = 0, cap_ = 0;
 public:
  IQMPS() {}
  IQMPS(int L, int D, int chi_cap);                       // empty device MPS
  // host data: bond_dims[L+1], charges (concatenated over bonds), tensors (concatenated, row-major [l][s][r])
  // llim / rlim: ITensor's orthogonality limits of the data (sites <= llim left-orthonormal, sites >= rlim right-orthonormal);
  // anything but (0, 2) is gauged to site 1 on the device after the upload (llim = 0, rlim = L+1: nothing assumed)
  IQMPS(int L, int D, int chi_cap, const std::vector<int>& bond_dims, const std::vector<int>& charges,
        const std::vector<Cplx>& tensors, int llim = 0, int rlim = 2);
  IQMPS(const IQMPS& o);                                  // deep copy on the device (ITensor copies are values)
  IQMPS& operator=(const IQMPS& o);
  IQMPS(IQMPS&&) = default;
  IQMPS& operator=(IQMPS&&) = default;
  explicit operator bool() const { return bool(p_); }
  int N() const { return L_; }
  int D() const { return D_; }
  int capacity() const { return cap_; }
  ocmps_mps* handle() const { return p_ ? p_->h : nullptr; }
  std::vector<int> bondDims() const;
  void toHost(std::vector<int>& bond_dims, std::vector<int>& charges, std::vector<Cplx>& tensors, int* llim = nullptr, int* rlim = nullptr) const;
  IQMPS withCapacity(int chi_cap) const;                  // re-homes the state in buffers of another capacity
};
inline LinkDim linkInd(const IQMPS& psi, int b) { return LinkDim{(long)psi.bondDims().at(b)}; }   // main/AnalyzeBondDim.cpp:140

// ---- IQMPO: the only MPO on the hot path is the propagator derivative K = sum_j 1/2 n_j(n_j-1) ----
class IQMPO {
 public:
  enum Kind { None, PropagatorDerivative };
  Kind kind = None;
  // the stepper that built this MPO (BH_tDMRG::propagatorDeriv): keeps its device gates alive and lets
  // exactApplyMPO(stepper.propagatorDeriv(u), psi, stepper.getArgs()) find the matching engine object directly
  std::shared_ptr<void> owner;
  ocmps_stepper* stepper = nullptr;
  double cutoff = -1.0;
  int maxm = 0;
  IQMPO() {}
  explicit IQMPO(Kind k) : kind(k) {}
};

Real norm(const IQMPS& psi);                                           // src/OptimalControl.cpp:257
Cplx overlapC(const IQMPS& a, const IQMPS& b);                         // <a|b>
Cplx overlapC(const IQMPS& a, const IQMPO& K, const IQMPS& b);         // <a|K|b>
void overlap(const IQMPS& a, const IQMPS& b, Real& re, Real& im);      // src/OptimalControl.cpp:450
// exactApplyMPO needs the stepper's truncation parameters; BH_tDMRG registers itself as the provider
IQMPS exactApplyMPO(const IQMPO& K, const IQMPS& psi, const Args& args);

}  // namespace itensor
#endif
