// Minimal stand-in for the IPOPT headers BH_nlp includes (IpTNLP.hpp & friends): just the types of the
// Ipopt::TNLP callback signatures, so that BH_nlp compiles and can be driven by a test harness or by any
// optimiser that speaks the TNLP interface.  With a real IPOPT installation put its include directory first.
#ifndef OCMPS_IPOPT_SHIM_IPTNLP_HPP
#define OCMPS_IPOPT_SHIM_IPTNLP_HPP
namespace Ipopt {
typedef int Index;
typedef double Number;
enum SolverReturn { SUCCESS, MAXITER_EXCEEDED, CPUTIME_EXCEEDED, STOP_AT_TINY_STEP, STOP_AT_ACCEPTABLE_POINT, LOCAL_INFEASIBILITY,
                    USER_REQUESTED_STOP, FEASIBLE_POINT_FOUND, DIVERGING_ITERATES, RESTORATION_FAILURE, ERROR_IN_STEP_COMPUTATION,
                    INVALID_NUMBER_DETECTED, TOO_FEW_DEGREES_OF_FREEDOM, INVALID_OPTION, OUT_OF_MEMORY, INTERNAL_ERROR, UNASSIGNED };
enum AlgorithmMode { RegularMode = 0, RestorationPhaseMode = 1 };
class IpoptData;
class IpoptCalculatedQuantities;
class TNLP {
 public:
  enum IndexStyleEnum { C_STYLE = 0, FORTRAN_STYLE = 1 };
  virtual ~TNLP() {}
  virtual bool get_nlp_info(Index& n, Index& m, Index& nnz_jac_g, Index& nnz_h_lag, IndexStyleEnum& index_style) = 0;
  virtual bool get_bounds_info(Index n, Number* x_l, Number* x_u, Index m, Number* g_l, Number* g_u) = 0;
  virtual bool get_starting_point(Index n, bool init_x, Number* x, bool init_z, Number* z_L, Number* z_U, Index m, bool init_lambda,
                                  Number* lambda) = 0;
  virtual bool eval_f(Index n, const Number* x, bool new_x, Number& obj_value) = 0;
  virtual bool eval_grad_f(Index n, const Number* x, bool new_x, Number* grad_f) = 0;
  virtual bool eval_g(Index n, const Number* x, bool new_x, Index m, Number* g) = 0;
  virtual bool eval_jac_g(Index n, const Number* x, bool new_x, Index m, Index nele_jac, Index* iRow, Index* jCol, Number* values) = 0;
  virtual bool eval_h(Index n, const Number* x, bool new_x, Number obj_factor, Index m, const Number* lambda, bool new_lambda,
                      Index nele_hess, Index* iRow, Index* jCol, Number* values) { return false; }
  virtual void finalize_solution(SolverReturn status, Index n, const Number* x, const Number* z_L, const Number* z_U, Index m,
                                 const Number* g, const Number* lambda, Number obj_value, const IpoptData* ip_data,
                                 IpoptCalculatedQuantities* ip_cq) = 0;
  virtual bool intermediate_callback(AlgorithmMode mode, Index iter, Number obj_value, Number inf_pr, Number inf_du, Number mu,
                                     Number d_norm, Number regularization_size, Number alpha_du, Number alpha_pr, Index ls_trials,
                                     const IpoptData* ip_data, IpoptCalculatedQuantities* ip_cq) { return true; }
};
}  // namespace Ipopt
#endif
