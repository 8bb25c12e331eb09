// Implementation of the itensor compatibility shim over libocmps.
#include "itensor/all.h"

#include <cstdlib>
#include <mutex>

namespace itensor {

ocmps_ctx* default_context() {
  static ocmps_ctx* ctx = nullptr;
  static std::once_flag once;
  std::call_once(once, [] {
    const char* e = std::getenv("OCMPS_DEVICE");
    ocmps_check(ocmps_ctx_create(e ? std::atoi(e) : 0, &ctx), "ocmps_ctx_create");
  });
  return ctx;
}

IQMPS::IQMPS(int L, int D, int chi_cap) : p_(std::make_shared<Holder>()), L_(L), D_(D), cap_(chi_cap) {
  ocmps_check(ocmps_mps_create(default_context(), L, D, chi_cap, &p_->h), "ocmps_mps_create");
}

IQMPS::IQMPS(int L, int D, int chi_cap, const std::vector<int>& bond_dims, const std::vector<int>& charges,
             const std::vector<Cplx>& tensors, int llim, int rlim)
    : IQMPS(L, D, chi_cap) {
  ocmps_check(ocmps_mps_upload(p_->h, bond_dims.data(), charges.data(), reinterpret_cast<const double*>(tensors.data()), llim, rlim),
              "ocmps_mps_upload");
  if (llim != 0 || rlim != 2) ocmps_check(ocmps_mps_position1(p_->h), "ocmps_mps_position1");
}

IQMPS::IQMPS(const IQMPS& o) : L_(o.L_), D_(o.D_), cap_(o.cap_) {
  if (o.p_) {
    p_ = std::make_shared<Holder>();
    ocmps_check(ocmps_mps_create(default_context(), L_, D_, cap_, &p_->h), "ocmps_mps_create");
    ocmps_check(ocmps_mps_copy(p_->h, o.p_->h), "ocmps_mps_copy");
  }
}

IQMPS& IQMPS::operator=(const IQMPS& o) {
  if (this != &o) {
    IQMPS tmp(o);
    std::swap(p_, tmp.p_);
    L_ = o.L_; D_ = o.D_; cap_ = o.cap_;
  }
  return *this;
}

std::vector<int> IQMPS::bondDims() const {
  std::vector<int> d(L_ + 1, 0);
  if (p_) ocmps_check(ocmps_mps_bond_dims(p_->h, d.data()), "ocmps_mps_bond_dims");
  return d;
}

void IQMPS::toHost(std::vector<int>& bond_dims, std::vector<int>& charges, std::vector<Cplx>& tensors, int* llim, int* rlim) const {
  long long ne = 0, nq = 0;
  ocmps_check(ocmps_mps_sizes(p_->h, &ne, &nq), "ocmps_mps_sizes");
  bond_dims.assign(L_ + 1, 0);
  charges.assign(nq, 0);
  tensors.assign(ne, Cplx(0.0, 0.0));
  int ll = 0, rl = 0;
  ocmps_check(ocmps_mps_download(p_->h, bond_dims.data(), charges.data(), reinterpret_cast<double*>(tensors.data()), &ll, &rl),
              "ocmps_mps_download");
  if (llim) *llim = ll;
  if (rlim) *rlim = rl;
}

IQMPS IQMPS::withCapacity(int chi_cap) const {
  if (chi_cap == cap_) return *this;
  std::vector<int> d, q;
  std::vector<Cplx> t;
  int ll = 0, rl = 2;
  toHost(d, q, t, &ll, &rl);
  return IQMPS(L_, D_, chi_cap, d, q, t, ll, rl);
}

Real norm(const IQMPS& psi) {
  double out = 0.0;
  ocmps_check(ocmps_mps_norm(psi.handle(), &out), "ocmps_mps_norm");
  return out;
}

Cplx overlapC(const IQMPS& a, const IQMPS& b) {
  double o[2];
  ocmps_check(ocmps_overlap(a.handle(), b.handle(), o), "ocmps_overlap");
  return Cplx(o[0], o[1]);
}

Cplx overlapC(const IQMPS& a, const IQMPO& K, const IQMPS& b) {
  if (K.kind != IQMPO::PropagatorDerivative) throw std::invalid_argument("overlapC: only the propagator derivative MPO is supported");
  double o[2];
  ocmps_check(ocmps_overlap_K(a.handle(), b.handle(), o), "ocmps_overlap_K");
  return Cplx(o[0], o[1]);
}

void overlap(const IQMPS& a, const IQMPS& b, Real& re, Real& im) {
  const Cplx z = overlapC(a, b);
  re = z.real();
  im = z.imag();
}

}  // namespace itensor

// ---- InputGroup ----
#include <cctype>
#include <fstream>
#include <sstream>
namespace itensor {
InputGroup::InputGroup(const std::string& filename, const std::string& groupname) : file_(filename), group_(groupname) {
  std::ifstream f(filename);
  if (!f.is_open()) throw std::runtime_error("InputGroup: cannot open " + filename);
  std::stringstream ss;
  ss << f.rdbuf();
  std::string txt = ss.str();
  for (size_t i = 0; i < txt.size(); ++i)                    // strip comments: '#' or '//' to the end of the line
    if (txt[i] == '#' || (txt[i] == '/' && i + 1 < txt.size() && txt[i + 1] == '/')) { while (i < txt.size() && txt[i] != '\n') txt[i++] = ' '; }
  size_t pos = 0;
  bool found = false;
  while ((pos = txt.find(groupname, pos)) != std::string::npos) {
    const bool left_ok = pos == 0 || std::isspace((unsigned char)txt[pos - 1]);
    size_t q = pos + groupname.size();
    while (q < txt.size() && std::isspace((unsigned char)txt[q])) ++q;
    if (left_ok && q < txt.size() && txt[q] == '{') { pos = q + 1; found = true; break; }
    pos += groupname.size();
  }
  if (!found) throw std::runtime_error("InputGroup: group '" + groupname + "' not found in " + filename);
  const size_t end = txt.find('}', pos);
  if (end == std::string::npos) throw std::runtime_error("InputGroup: missing '}' in " + filename);
  std::stringstream body(txt.substr(pos, end - pos));
  std::string line;
  while (std::getline(body, line)) {
    const size_t eq = line.find('=');
    if (eq == std::string::npos) continue;
    auto trim = [](std::string v) {
      size_t a = 0, b = v.size();
      while (a < b && std::isspace((unsigned char)v[a])) ++a;
      while (b > a && std::isspace((unsigned char)v[b - 1])) --b;
      return v.substr(a, b - a);
    };
    const std::string k = trim(line.substr(0, eq)), v = trim(line.substr(eq + 1));
    if (!k.empty() && !v.empty()) kv_[k] = v;
  }
}
double InputGroup::getReal(const std::string& k) const {
  const std::string* v = find(k);
  if (!v) throw std::runtime_error("InputGroup: mandatory key '" + k + "' missing in group '" + group_ + "' of " + file_);
  return std::stod(*v);
}
int InputGroup::getInt(const std::string& k) const {
  const std::string* v = find(k);
  if (!v) throw std::runtime_error("InputGroup: mandatory key '" + k + "' missing in group '" + group_ + "' of " + file_);
  return (int)std::stol(*v);
}
bool InputGroup::getYesNo(const std::string& k) const {
  const std::string* v = find(k);
  if (!v) throw std::runtime_error("InputGroup: mandatory key '" + k + "' missing in group '" + group_ + "' of " + file_);
  std::string t = *v;
  for (char& c : t) c = (char)std::tolower((unsigned char)c);
  if (t == "yes" || t == "y" || t == "true" || t == "1") return true;
  if (t == "no" || t == "n" || t == "false" || t == "0") return false;
  throw std::runtime_error("InputGroup: key '" + k + "' is neither yes nor no");
}
std::string InputGroup::getString(const std::string& k) const {
  const std::string* v = find(k);
  if (!v) throw std::runtime_error("InputGroup: mandatory key '" + k + "' missing");
  return *v;
}
}  // namespace itensor

