// ref_functors.cpp -- evaluates the REFERENCE'S OWN Ceres cost functors (/root/reference/src/lidarFeaturePointsFunction.hpp,
// included from where it lies) on dual numbers.  TEST INFRASTRUCTURE: it pins SURVEY section 8 rows b4 (LidarEdgeFactor
// :243-293), b5 (LidarPlaneFactor :143-196), c6 (LidarPlaneNormFactor :199-240) and f3's front_end_residual (:21-58): the
// residual formulas are the reference's text, differentiated here the way ceres::AutoDiffCostFunction does it (forward
// mode, one Jet carrying d/d{qx,qy,qz,qw,tx,ty,tz}).  Eigen and Ceres are not installed: the functors compile against the
// stand-ins of oracle/shims/ (Eigen 3.3's formulas for the 3-vector / quaternion pieces they use).
// Output: oracle/_ref/libref_functors.so (git-ignored).
#include <cmath>

struct Jet7 {
  double a;
  double v[7];
  Jet7() : a(0), v{0, 0, 0, 0, 0, 0, 0} {}
  Jet7(double s) : a(s), v{0, 0, 0, 0, 0, 0, 0} {}  // NOLINT: T(double) conversions are what the functors rely on
  Jet7(double s, int k) : a(s), v{0, 0, 0, 0, 0, 0, 0} { v[k] = 1.0; }
};
#define JET_LOOP for (int i = 0; i < 7; ++i)
inline Jet7 operator+(const Jet7& f, const Jet7& g) { Jet7 h; h.a = f.a + g.a; JET_LOOP h.v[i] = f.v[i] + g.v[i]; return h; }
inline Jet7 operator-(const Jet7& f, const Jet7& g) { Jet7 h; h.a = f.a - g.a; JET_LOOP h.v[i] = f.v[i] - g.v[i]; return h; }
inline Jet7 operator-(const Jet7& f) { Jet7 h; h.a = -f.a; JET_LOOP h.v[i] = -f.v[i]; return h; }
inline Jet7 operator*(const Jet7& f, const Jet7& g) { Jet7 h; h.a = f.a * g.a; JET_LOOP h.v[i] = f.a * g.v[i] + f.v[i] * g.a; return h; }
inline Jet7 operator/(const Jet7& f, const Jet7& g) {
  Jet7 h; const double inv = 1.0 / g.a; h.a = f.a * inv;
  JET_LOOP h.v[i] = (f.v[i] - h.a * g.v[i]) * inv;
  return h;
}
inline bool operator<(const Jet7& f, const Jet7& g) { return f.a < g.a; }
inline bool operator>(const Jet7& f, const Jet7& g) { return f.a > g.a; }
inline bool operator>=(const Jet7& f, const Jet7& g) { return f.a >= g.a; }
namespace std {
inline Jet7 sqrt(const Jet7& f) { Jet7 h; h.a = ::sqrt(f.a); const double d = 0.5 / h.a; JET_LOOP h.v[i] = f.v[i] * d; return h; }
inline Jet7 sin(const Jet7& f) { Jet7 h; h.a = ::sin(f.a); const double d = ::cos(f.a); JET_LOOP h.v[i] = f.v[i] * d; return h; }
inline Jet7 acos(const Jet7& f) { Jet7 h; h.a = ::acos(f.a); const double d = -1.0 / ::sqrt(1.0 - f.a * f.a); JET_LOOP h.v[i] = f.v[i] * d; return h; }
inline Jet7 abs(const Jet7& f) { return f.a < 0.0 ? -f : f; }
}  // namespace std

#include <lidarFeaturePointsFunction.hpp>

#define REF_API extern "C" __attribute__((visibility("default")))

namespace {
template <class F, int R>
void run(const F& f, const double* q, const double* t, double* r, double* J) {
  Jet7 jq[4], jt[3], jr[R];
  for (int i = 0; i < 4; ++i) jq[i] = Jet7(q[i], i);
  for (int i = 0; i < 3; ++i) jt[i] = Jet7(t[i], 4 + i);
  f(jq, jt, jr);
  for (int k = 0; k < R; ++k) {
    r[k] = jr[k].a;
    for (int i = 0; i < 7; ++i) J[k * 7 + i] = jr[k].v[i];
  }
  // residuals from the plain-double instantiation (what Ceres uses when it evaluates the cost without Jacobians); the
  // Jet's scalar part differs from it by at most an ulp (the Jet division multiplies by the reciprocal, as ceres::Jet's)
  double rd[R];
  f(q, t, rd);
  for (int k = 0; k < R; ++k) r[k] = std::fabs(rd[k] - jr[k].a) <= 1e-12 * (1.0 + std::fabs(rd[k])) ? rd[k] : NAN;
}
Eigen::Vector3d v3(const double* p) { return Eigen::Vector3d(p[0], p[1], p[2]); }
}  // namespace

// type: 1 LidarEdgeFactor(curr, a, b, s)          -> 3 residuals      laserMapping.cpp:718, laserOdometry.cpp:556-559
//       2 LidarPlaneNormFactor(curr, n, d)         -> 1 residual       laserMapping.cpp:791, mapOptimization.cpp:425
//       3 front_end_residual(src, dst)             -> 3 residuals      intensity_feature_tracker.cpp:900
//       4 LidarPlaneFactor(curr, j, l, m, s)       -> 1 residual       laserOdometry.cpp:679-682
// q = (x, y, z, w), t = (x, y, z); r: up to 3 residuals; J: rows x 7 ambient Jacobian d r / d (q, t), row-major.
// Returns the number of residual rows.
REF_API int ref_functor_eval(int type, const double* p, const double* a, const double* b, const double* c, double s, const double* q,
                             const double* t, double* r, double* J) {
  switch (type) {
    case 1: run<LidarEdgeFactor, 3>(LidarEdgeFactor(v3(p), v3(a), v3(b), s), q, t, r, J); return 3;
    case 2: run<LidarPlaneNormFactor, 1>(LidarPlaneNormFactor(v3(p), v3(a), b[0]), q, t, r, J); return 1;
    case 3: run<front_end_residual, 3>(front_end_residual(v3(p), v3(a)), q, t, r, J); return 3;
    case 4: run<LidarPlaneFactor, 1>(LidarPlaneFactor(v3(p), v3(a), v3(b), v3(c), s), q, t, r, J); return 1;
  }
  return 0;
}
