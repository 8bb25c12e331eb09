// u(t_i) = u0(t_i) + S(t_i) * sum_n c_n f_n(t_i): GROUP parametrisation of the control
// (interface of the reference's include/ControlBasis.hpp:30-39; flat row-major storage inside).
#ifndef OCMPS_CONTROLBASIS_HPP
#define OCMPS_CONTROLBASIS_HPP
#include <cassert>
#include <cstddef>
#include <vector>

typedef std::vector<double> stdvec;
typedef std::vector<std::vector<double>> rowmat;

class ControlBasis {
  size_t N_ = 0, M_ = 0;
  stdvec u0_, S_, f_, jac_, current_;      // f_, jac_: N x M row-major
 public:
  ControlBasis() {}
  ControlBasis(stdvec& u0, stdvec& S, rowmat& f);
  size_t getM() const { return M_; }
  size_t getN() const { return N_; }
  stdvec convertControl(const stdvec& control, const bool new_control = true);
  stdvec convertGradient(const stdvec& gradu) const;
  rowmat convertHessian(const rowmat& Hessu) const;
  rowmat getControlJacobian() const;
};
#endif
