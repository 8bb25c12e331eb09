// GROUP control basis: host-side O(N*M) projections (the reference keeps them in scalar loops as well,
// src/ControlBasis.cpp:49-119); summation orders follow the reference so results agree to the last bit.
#include "ControlBasis.hpp"

ControlBasis::ControlBasis(stdvec& u0, stdvec& S, rowmat& f) : N_(u0.size()), M_(f.front().size()), u0_(u0), S_(S) {
  f_.resize(N_ * M_);
  jac_.resize(N_ * M_);
  for (size_t i = 0; i < N_; ++i)
    for (size_t n = 0; n < M_; ++n) {
      f_[i * M_ + n] = f[i][n];
      jac_[i * M_ + n] = f[i][n] * S[i];        // du_i / dc_n
    }
  current_ = u0_;
}

stdvec ControlBasis::convertControl(const stdvec& control, const bool new_control) {
  if (new_control) {                             // otherwise the cached control is returned
    assert(control.size() == M_);
    stdvec u(u0_);
    for (size_t i = 0; i < N_; ++i) {
      double acc = 0.0;
      for (size_t n = 0; n < M_; ++n) acc += f_[i * M_ + n] * control[n];
      u[i] += S_[i] * acc;
    }
    current_.swap(u);
  }
  return current_;
}

stdvec ControlBasis::convertGradient(const stdvec& gradu) const {
  assert(gradu.size() == N_);
  stdvec gc(M_, 0.0);
  for (size_t n = 0; n < M_; ++n) {
    double acc = 0.0;
    for (size_t i = 0; i < N_; ++i) acc += S_[i] * gradu[i] * f_[i * M_ + n];
    gc[n] = acc;
  }
  return gc;
}

rowmat ControlBasis::convertHessian(const rowmat& Hessu) const {
  assert(Hessu.size() == N_ && Hessu.front().size() == N_);
  rowmat Hc(M_, stdvec(M_, 0.0));
  stdvec Hv(N_);
  for (size_t j = 0; j < M_; ++j) {
    for (size_t k = 0; k < N_; ++k) {            // Hv = H . v_j with v_j = column j of the Jacobian
      double acc = 0.0;
      for (size_t l = 0; l < N_; ++l) acc += Hessu[k][l] * jac_[l * M_ + j];
      Hv[k] = acc;
    }
    for (size_t i = 0; i <= j; ++i) {            // upper triangle, mirrored
      double acc = 0.0;
      for (size_t k = 0; k < N_; ++k) acc += jac_[k * M_ + i] * Hv[k];
      Hc[i][j] = acc;
      Hc[j][i] = acc;
    }
  }
  return Hc;
}

rowmat ControlBasis::getControlJacobian() const {
  rowmat J(N_, stdvec(M_));
  for (size_t i = 0; i < N_; ++i)
    for (size_t n = 0; n < M_; ++n) J[i][n] = jac_[i * M_ + n];
  return J;
}
