// ceres/ceres.h -- stand-in for the four Ceres base classes the reference's factors derive from (interfaces only:
// no solver).  Ceres is not installed in this image.  TEST INFRASTRUCTURE ONLY; our own code.
#pragma once
#include <cmath>
#include <string>
#include <vector>
namespace ceres {
class CostFunction {
 public:
  virtual ~CostFunction() {}
  virtual bool Evaluate(double const* const* parameters, double* residuals, double** jacobians) const = 0;
  const std::vector<int>& parameter_block_sizes() const { return sizes_; }
  int num_residuals() const { return nres_; }
 protected:
  std::vector<int>* mutable_parameter_block_sizes() { return &sizes_; }
  void set_num_residuals(int n) { nres_ = n; }
 private:
  std::vector<int> sizes_;
  int nres_ = 0;
};
template <int kNumResiduals, int... Ns> class SizedCostFunction : public CostFunction {
 public:
  SizedCostFunction() { set_num_residuals(kNumResiduals); *mutable_parameter_block_sizes() = std::vector<int>{Ns...}; }
};
class LossFunction {
 public:
  virtual ~LossFunction() {}
  virtual void Evaluate(double sq_norm, double out[3]) const = 0;
};
// rho(s) = b log(1 + s / b), b = a^2 (Ceres' documented CauchyLoss)
class CauchyLoss : public LossFunction {
 public:
  explicit CauchyLoss(double a) : b_(a * a), c_(1 / b_) {}
  void Evaluate(double s, double rho[3]) const override {
    const double sum = 1 + s * c_, inv = 1 / sum;
    rho[0] = b_ * std::log(sum); rho[1] = inv; rho[2] = -c_ * (inv * inv);
  }
 private:
  const double b_, c_;
};
// named by initial/initial_sfm.h (ReprojectionError3D::Create); never instantiated here
template <class F, int... Ns> class AutoDiffCostFunction : public CostFunction {
 public:
  explicit AutoDiffCostFunction(F*) {}
  bool Evaluate(double const* const*, double*, double**) const override { return false; }
};
class LocalParameterization {
 public:
  virtual ~LocalParameterization() {}
  virtual bool Plus(const double* x, const double* delta, double* x_plus_delta) const = 0;
  virtual bool ComputeJacobian(const double* x, double* jacobian) const = 0;
  virtual int GlobalSize() const = 0;
  virtual int LocalSize() const = 0;
};
// ---- the Problem / Solve surface Estimator::optimization() uses (estimator.cpp:663-811) -----------------------------------------
// No solver lives here.  Problem records what the reference hands to Ceres; Solve() forwards to a hook the test driver
// installs (it injects a solution computed elsewhere), so that the reference code AROUND the solve -- vector2double,
// the problem construction, double2vector, the marginalization glue -- runs unmodified.
enum LinearSolverType { DENSE_QR, DENSE_SCHUR, SPARSE_SCHUR, SPARSE_NORMAL_CHOLESKY, DENSE_NORMAL_CHOLESKY, ITERATIVE_SCHUR, CGNR };
enum TrustRegionStrategyType { LEVENBERG_MARQUARDT, DOGLEG };
class HuberLoss : public LossFunction {
 public:
  explicit HuberLoss(double a) : a_(a), b_(a * a) {}
  void Evaluate(double s, double rho[3]) const override {
    if (s > b_) { const double r = std::sqrt(s); rho[0] = 2 * a_ * r - b_; rho[1] = a_ / r; rho[2] = -rho[1] / (2 * s); }
    else { rho[0] = s; rho[1] = 1; rho[2] = 0; }
  }
 private:
  const double a_, b_;
};
class Problem {
 public:
  struct ParameterBlock { double* ptr; int size; LocalParameterization* local; bool constant; };
  struct ResidualBlock { CostFunction* cost; LossFunction* loss; std::vector<double*> blocks; };
  std::vector<ParameterBlock> parameter_blocks;
  std::vector<ResidualBlock> residual_blocks;
  void AddParameterBlock(double* p, int size, LocalParameterization* lp = nullptr) { parameter_blocks.push_back({p, size, lp, false}); }
  void SetParameterBlockConstant(double* p) { for (auto& b : parameter_blocks) if (b.ptr == p) b.constant = true; }
  template <class... Ps> void AddResidualBlock(CostFunction* c, LossFunction* l, Ps... ps) { residual_blocks.push_back({c, l, std::vector<double*>{ps...}}); }
  void AddResidualBlock(CostFunction* c, LossFunction* l, const std::vector<double*>& ps) { residual_blocks.push_back({c, l, ps}); }
};
class Solver {
 public:
  struct Options {
    LinearSolverType linear_solver_type = SPARSE_NORMAL_CHOLESKY;
    TrustRegionStrategyType trust_region_strategy_type = LEVENBERG_MARQUARDT;
    int max_num_iterations = 50, num_threads = 1;
    double max_solver_time_in_seconds = 1e9;
    bool minimizer_progress_to_stdout = false, use_explicit_schur_complement = false;
  };
  struct Summary {
    std::vector<int> iterations;
    std::string BriefReport() const { return "ceres stand-in: no solver"; }
    std::string FullReport() const { return BriefReport(); }
  };
};
typedef void (*SolveHook)(const Solver::Options&, Problem*, Solver::Summary*);
inline SolveHook& solve_hook() { static SolveHook h = nullptr; return h; }
inline void Solve(const Solver::Options& o, Problem* p, Solver::Summary* s) { if (solve_hook()) solve_hook()(o, p, s); }
}  // namespace ceres
