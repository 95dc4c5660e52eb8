// Stand-in for <ceres/ceres.h> (Ceres is not installed here): the reference's functors only NAME ceres::CostFunction and
// ceres::AutoDiffCostFunction in their static Create() helpers (lidarFeaturePointsFunction.hpp:51,91,131,187,231,285);
// the derivatives for the pin come from oracle/ref_functors.cpp's own dual numbers.  TEST INFRASTRUCTURE.
#pragma once
namespace ceres {
class CostFunction {
 public:
  virtual ~CostFunction() {}
};
template <typename Functor, int kNumResiduals, int N0 = 0, int N1 = 0, int N2 = 0>
class AutoDiffCostFunction : public CostFunction {
 public:
  explicit AutoDiffCostFunction(Functor* f) : functor_(f) {}
  ~AutoDiffCostFunction() override { delete functor_; }
  Functor* functor_;
};
}  // namespace ceres
