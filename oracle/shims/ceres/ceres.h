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

// The problem-building API as far as laserOdometry.cpp:417-713 uses it.  This Problem does not optimise anything: it
// RECORDS the residual blocks it is handed (oracle/ref_laserodom.cpp reads the functors back to compare the reference's
// data association with the oracle's), and Solve() leaves the parameters where they are.
class LossFunction {
 public:
  virtual ~LossFunction() {}
};
class HuberLoss : public LossFunction {
 public:
  explicit HuberLoss(double a) : a_(a) {}
  double a_;
};
class LocalParameterization {
 public:
  virtual ~LocalParameterization() {}
};
class EigenQuaternionParameterization : public LocalParameterization {};
class Problem {
 public:
  struct Options {};
  Problem() {}
  explicit Problem(const Options&) {}
  ~Problem();
  void AddParameterBlock(double*, int) {}
  void AddParameterBlock(double*, int, LocalParameterization* p) { params_ = p; }
  void AddResidualBlock(CostFunction* f, LossFunction* l, double*, double*) {
    loss_ = l;
    if (sink) sink(f, sink_arg);
    blocks_++;
    delete f;
  }
  static void (*sink)(CostFunction*, void*);
  static void* sink_arg;
  int blocks_ = 0;
  LossFunction* loss_ = nullptr;
  LocalParameterization* params_ = nullptr;
};
inline Problem::~Problem() {
  delete loss_;
  delete params_;
}
enum LinearSolverType { DENSE_QR };
struct Solver {
  struct Options {
    LinearSolverType linear_solver_type = DENSE_QR;
    int max_num_iterations = 50;
    bool minimizer_progress_to_stdout = false;
    bool check_gradients = false;
    double gradient_check_relative_precision = 1e-8;
  };
  struct Summary {};
};
inline void Solve(const Solver::Options&, Problem*, Solver::Summary*) {}
}  // namespace ceres
