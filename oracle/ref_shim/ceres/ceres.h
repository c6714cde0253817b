// ceres/ceres.h -- stand-in for the four Ceres base classes the reference's factors derive from (interfaces only:
// no solver).  Ceres is not installed in this image.  TEST INFRASTRUCTURE ONLY; our own code.
#pragma once
#include <cmath>
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
}  // namespace ceres
