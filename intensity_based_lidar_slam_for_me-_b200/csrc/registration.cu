// registration.cu -- K2 (per-correspondence line/plane fit, fused behind the k-NN) and K3 (residual + 6-DoF
// Jacobian + J^T J / J^T r reduction with the Levenberg-Marquardt update folded into the last block).
//
// Replaces the association + ceres::Solve block of laserMapping.cpp:640-861 and mapOptimization.cpp:377-450:
//   associate_kernel : pointAssociateToMap (double math, float store) -> exact 5-NN in the voxel hash ->
//                      gate d2[4] < 1 -> 3x3 scatter-matrix eigen (corner) / 5x3 pivoted-QR plane (surf),
//                      all in registers; one factor slot per stack point.
//   eval_kernel      : LidarEdgeFactor / LidarPlaneNormFactor residuals with closed-form tangent Jacobians
//                      (hpp:199-293 + EigenQuaternionParameterization), HuberLoss(0.1) corrector, warp-shuffle +
//                      block reduction of cost/JtJ/Jtr, deterministic cross-block sum by the last block, which
//                      then advances the trust-region state machine (Ceres 1.14 TrustRegionMinimizer +
//                      LevenbergMarquardtStrategy restated) and writes the next candidate pose to HBM.
// A registration is therefore a fixed sequence of launches with no host round trip until the final pose.
#include "ilsm_host.hpp"

namespace ilsm {

// ---------------------------------------------------------------------------------------------------
// fits
// ---------------------------------------------------------------------------------------------------
__device__ __forceinline__ void jacobi_rot(double& app, double& aqq, double& apq, double& arp, double& arq, double& v0p,
                                           double& v0q, double& v1p, double& v1q, double& v2p, double& v2q) {
  if (apq == 0.0) return;
  double theta = (aqq - app) / (2.0 * apq);
  double t = (theta >= 0.0 ? 1.0 : -1.0) / (fabs(theta) + sqrt(theta * theta + 1.0));
  double c = 1.0 / sqrt(t * t + 1.0), s = t * c;
  app -= t * apq;
  aqq += t * apq;
  apq = 0.0;
  double x = arp, y = arq;
  arp = c * x - s * y;
  arq = s * x + c * y;
  x = v0p, y = v0q, v0p = c * x - s * y, v0q = s * x + c * y;
  x = v1p, y = v1q, v1p = c * x - s * y, v1q = s * x + c * y;
  x = v2p, y = v2q, v2p = c * x - s * y, v2q = s * x + c * y;
}

// laserMapping.cpp:681-722.  nb = the 5 neighbours (float map points widened to double).
__device__ __forceinline__ bool fit_line(const float (&nb)[5][3], double ratio, double (&pa)[3], double (&pb)[3]) {
  double cx = 0, cy = 0, cz = 0;
#pragma unroll
  for (int j = 0; j < 5; ++j) cx = cx + (double)nb[j][0], cy = cy + (double)nb[j][1], cz = cz + (double)nb[j][2];
  cx = cx / 5.0, cy = cy / 5.0, cz = cz / 5.0;
  double a00 = 0, a01 = 0, a02 = 0, a11 = 0, a12 = 0, a22 = 0;
#pragma unroll
  for (int j = 0; j < 5; ++j) {
    double zx = (double)nb[j][0] - cx, zy = (double)nb[j][1] - cy, zz = (double)nb[j][2] - cz;
    a00 = a00 + zx * zx, a01 = a01 + zx * zy, a02 = a02 + zx * zz;
    a11 = a11 + zy * zy, a12 = a12 + zy * zz, a22 = a22 + zz * zz;
  }
  double v00 = 1, v01 = 0, v02 = 0, v10 = 0, v11 = 1, v12 = 0, v20 = 0, v21 = 0, v22 = 1;
#pragma unroll 1
  for (int sweep = 0; sweep < 30; ++sweep) {
    double off = a01 * a01 + a02 * a02 + a12 * a12;
    double dg = a00 * a00 + a11 * a11 + a22 * a22;
    if (off <= 1e-32 * dg || off == 0.0) break;
    jacobi_rot(a00, a11, a01, a02, a12, v00, v01, v10, v11, v20, v21);  // (p,q)=(0,1), r=2
    jacobi_rot(a00, a22, a02, a01, a12, v00, v02, v10, v12, v20, v22);  // (0,2), r=1
    jacobi_rot(a11, a22, a12, a01, a02, v01, v02, v11, v12, v21, v22);  // (1,2), r=0
  }
  // largest eigenvalue / vector and the middle eigenvalue
  double lmax = a00, lmid, dx = v00, dy = v10, dz = v20;
  double o1 = a11, o2 = a22;
  if (a11 > lmax) lmax = a11, dx = v01, dy = v11, dz = v21, o1 = a00, o2 = a22;
  if (a22 > lmax) lmax = a22, dx = v02, dy = v12, dz = v22, o1 = a00, o2 = a11;
  lmid = o1 > o2 ? o1 : o2;
  if (!(lmax > ratio * lmid)) return false;
  pa[0] = 0.1 * dx + cx, pa[1] = 0.1 * dy + cy, pa[2] = 0.1 * dz + cz;
  pb[0] = -0.1 * dx + cx, pb[1] = -0.1 * dy + cy, pb[2] = -0.1 * dz + cz;
  return true;
}

// laserMapping.cpp:756-796 / mapOptimization.cpp:395-427: A(5x3) n = -1 by Householder QR with column pivoting.
__device__ __forceinline__ bool fit_plane(const float (&nb)[5][3], double tol, double (&nrm)[3], double& d_out) {
  double A[5][3], b[5];
#pragma unroll
  for (int i = 0; i < 5; ++i) {
    b[i] = -1.0;
#pragma unroll
    for (int c = 0; c < 3; ++c) A[i][c] = (double)nb[i][c];
  }
  int perm[3] = {0, 1, 2};
#pragma unroll
  for (int k = 0; k < 3; ++k) {
    double cn[3] = {0, 0, 0};
#pragma unroll
    for (int j = k; j < 3; ++j) {
#pragma unroll
      for (int i = k; i < 5; ++i) cn[j] += A[i][j] * A[i][j];
    }
    int piv = k;
    double best = cn[k];
#pragma unroll
    for (int j = k + 1; j < 3; ++j)
      if (cn[j] > best) best = cn[j], piv = j;
#pragma unroll
    for (int j = k + 1; j < 3; ++j) {
      if (piv == j) {
#pragma unroll
        for (int i = 0; i < 5; ++i) {
          double tmp = A[i][k];
          A[i][k] = A[i][j];
          A[i][j] = tmp;
        }
        int tp = perm[k];
        perm[k] = perm[j];
        perm[j] = tp;
      }
    }
    double nr = sqrt(best);
    if (nr == 0.0) continue;
    double alpha = A[k][k] > 0.0 ? -nr : nr;
    double v[5];
#pragma unroll
    for (int i = 0; i < 5; ++i) v[i] = i >= k ? A[i][k] : 0.0;
    v[k] -= alpha;
    double vtv = 0;
#pragma unroll
    for (int i = k; i < 5; ++i) vtv += v[i] * v[i];
    if (vtv == 0.0) continue;
#pragma unroll
    for (int j = k; j < 3; ++j) {
      double s = 0;
#pragma unroll
      for (int i = k; i < 5; ++i) s += v[i] * A[i][j];
      s = 2.0 * s / vtv;
#pragma unroll
      for (int i = k; i < 5; ++i) A[i][j] -= s * v[i];
    }
    double s = 0;
#pragma unroll
    for (int i = k; i < 5; ++i) s += v[i] * b[i];
    s = 2.0 * s / vtv;
#pragma unroll
    for (int i = k; i < 5; ++i) b[i] -= s * v[i];
  }
  double y[3];
#pragma unroll
  for (int k = 2; k >= 0; --k) {
    double s = b[k];
#pragma unroll
    for (int j = k + 1; j < 3; ++j) s -= A[k][j] * y[j];
    y[k] = A[k][k] != 0.0 ? s / A[k][k] : 0.0;
  }
  double n[3] = {0, 0, 0};
#pragma unroll
  for (int k = 0; k < 3; ++k) {
#pragma unroll
    for (int c = 0; c < 3; ++c)
      if (perm[k] == c) n[c] = y[k];
  }
  double nn = sqrt(n[0] * n[0] + n[1] * n[1] + n[2] * n[2]);
  double d = 1.0 / nn;
  n[0] /= nn, n[1] /= nn, n[2] /= nn;
  bool ok = nn > 0.0 && isfinite(d);
#pragma unroll
  for (int j = 0; j < 5; ++j)
    if (fabs(n[0] * (double)nb[j][0] + n[1] * (double)nb[j][1] + n[2] * (double)nb[j][2] + d) > tol) ok = false;
  nrm[0] = n[0], nrm[1] = n[1], nrm[2] = n[2];
  d_out = d;
  return ok;
}

struct FactorView {
  int* type;
  float4* p;
  double4* a;
  double4* b;
  int32_t* knn_idx;  // may be null
  float* knn_d2;     // may be null
};

struct AssocParams {
  float gate_sq;
  double line_ratio, plane_tol;
  int begin_solve, pass, max_iter;
  double huber_a;
};

__device__ void lm_begin(LmState* st, const AssocParams& prm) {
#pragma unroll
  for (int i = 0; i < 4; ++i) st->cq[i] = st->xq[i];
#pragma unroll
  for (int i = 0; i < 3; ++i) st->ct[i] = st->xt[i];
  st->status = 0;
  st->phase = 0;
  st->iteration = 0;
  st->max_iter = prm.max_iter;
  st->invalid_run = 0;
  st->reuse_diag = 0;
  st->n_success = st->n_unsuccess = st->n_evals = 0;
  st->n_edge = st->n_plane = 0;
  st->radius = 1e4;
  st->decrease_factor = 2.0;
  st->model_cost_change = 0.0;
  st->huber_a = prm.huber_a;
  st->pass = prm.pass;
  st->ticket = 0u;
}

template <int G>
__global__ void __launch_bounds__(256)
    associate_kernel(GridView gc, GridView gs, const float* __restrict__ corner, int nc, const float* __restrict__ surf,
                     int ns, int stride_f, LmState* st, AssocParams prm, FactorView fv) {
  if (prm.begin_solve && blockIdx.x == 0 && threadIdx.x == 0) lm_begin(st, prm);
  const int groups_per_block = blockDim.x / G;
  const int gid = blockIdx.x * groups_per_block + threadIdx.x / G;
  const unsigned lane = threadIdx.x % G;
  const unsigned wl = threadIdx.x & 31;
  const unsigned gmask = (G == 32) ? 0xffffffffu : (((1u << G) - 1u) << (wl - lane));
  if (gid >= nc + ns) return;
  const bool is_corner = gid < nc;
  const float* pp = is_corner ? corner + (size_t)gid * stride_f : surf + (size_t)(gid - nc) * stride_f;
  const float px = __ldg(pp), py = __ldg(pp + 1), pz = __ldg(pp + 2);
  double q[4] = {st->xq[0], st->xq[1], st->xq[2], st->xq[3]};
  // pointAssociateToMap: double math, float store
  D3 pw = quat_rotate(q, d3((double)px, (double)py, (double)pz));
  const float qx = __double2float_rn(dadd(pw.x, st->xt[0]));
  const float qy = __double2float_rn(dadd(pw.y, st->xt[1]));
  const float qz = __double2float_rn(dadd(pw.z, st->xt[2]));

  u64 res[5];
  if (is_corner)
    knn_search<5, G>(gc, qx, qy, qz, prm.gate_sq, lane, gmask, res);
  else
    knn_search<5, G>(gs, qx, qy, qz, prm.gate_sq, lane, gmask, res);

  const bool gate = res[4] != kSentinel && cand_d2(res[4]) < prm.gate_sq;
  int type = 0;
  double a[3] = {0, 0, 0}, b[3] = {0, 0, 0}, w = 0;
  if (gate) {  // group-uniform
    // lanes 0..4 gather one neighbour each (original-order array), then everything is shuffled to lane 0
    float4 me = make_float4(0.f, 0.f, 0.f, 0.f);
    {
      int my = 0;
#pragma unroll
      for (int k = 0; k < 5; ++k)
        if ((int)lane == k) my = cand_idx(res[k]);
      if (lane < 5) me = __ldg((is_corner ? gc.orig : gs.orig) + my);
    }
    float nb[5][3];
#pragma unroll
    for (int k = 0; k < 5; ++k) {
      nb[k][0] = __shfl_sync(gmask, me.x, k, G);
      nb[k][1] = __shfl_sync(gmask, me.y, k, G);
      nb[k][2] = __shfl_sync(gmask, me.z, k, G);
    }
    if (lane == 0) {
      if (is_corner) {
        if (fit_line(nb, prm.line_ratio, a, b)) type = 1;
      } else {
        if (fit_plane(nb, prm.plane_tol, a, w)) type = 2;
      }
    }
  }
  if (lane == 0) {
    fv.type[gid] = type;
    fv.p[gid] = make_float4(px, py, pz, 0.f);
    fv.a[gid] = make_double4(a[0], a[1], a[2], w);
    fv.b[gid] = make_double4(b[0], b[1], b[2], 0.0);
    if (fv.knn_idx) {
#pragma unroll
      for (int k = 0; k < 5; ++k) {
        bool have = res[k] != kSentinel;
        fv.knn_idx[(size_t)gid * 5 + k] = have ? cand_idx(res[k]) : -1;
        fv.knn_d2[(size_t)gid * 5 + k] = have ? cand_d2(res[k]) : __int_as_float(0x7f800000);
      }
    }
  }
}

// ---------------------------------------------------------------------------------------------------
// Levenberg-Marquardt state machine (one thread, state staged in shared memory)
// ---------------------------------------------------------------------------------------------------
constexpr int kNumSums = 30;  // cost, H[21], g[6], #edge, #plane
constexpr int kSumStride = 32;

__device__ __forceinline__ void quat_mul_d(const double a[4], const double b[4], double o[4]) {
  o[0] = a[3] * b[0] + a[0] * b[3] + a[1] * b[2] - a[2] * b[1];
  o[1] = a[3] * b[1] + a[1] * b[3] + a[2] * b[0] - a[0] * b[2];
  o[2] = a[3] * b[2] + a[2] * b[3] + a[0] * b[1] - a[1] * b[0];
  o[3] = a[3] * b[3] - a[0] * b[0] - a[1] * b[1] - a[2] * b[2];
}
// EigenQuaternionParameterization::Plus
__device__ void quat_plus_d(const double x[4], const double d[3], double o[4]) {
  double nd = sqrt(d[0] * d[0] + d[1] * d[1] + d[2] * d[2]);
  if (nd > 0.0) {
    double sbd = sin(nd) / nd;
    double dq[4] = {sbd * d[0], sbd * d[1], sbd * d[2], cos(nd)};
    quat_mul_d(dq, x, o);
  } else {
    o[0] = x[0], o[1] = x[1], o[2] = x[2], o[3] = x[3];
  }
}

__device__ bool chol_solve6_d(double (&A)[6][6], const double (&b)[6], double (&x)[6]) {
  // in-place lower Cholesky
  for (int i = 0; i < 6; ++i)
    for (int j = 0; j <= i; ++j) {
      double s = A[i][j];
      for (int k = 0; k < j; ++k) s -= A[i][k] * A[j][k];
      if (i == j) {
        if (!(s > 0.0)) return false;
        A[i][i] = sqrt(s);
      } else {
        A[i][j] = s / A[j][j];
      }
    }
  double y[6];
  for (int i = 0; i < 6; ++i) {
    double s = b[i];
    for (int k = 0; k < i; ++k) s -= A[i][k] * y[k];
    y[i] = s / A[i][i];
  }
  for (int i = 5; i >= 0; --i) {
    double s = y[i];
    for (int k = i + 1; k < 6; ++k) s -= A[k][i] * x[k];
    x[i] = s / A[i][i];
  }
  return true;
}

__device__ void lm_terminate(LmState* s, LmState* g, int code) {
  s->status = 1 + code;
  ilsm_solve_summary& r = g->report.pass[s->pass < ILSM_MAX_OUTER ? s->pass : ILSM_MAX_OUTER - 1];
  r.termination = code;
  r.iterations = s->iteration;
  r.num_successful_steps = s->n_success;
  r.num_unsuccessful_steps = s->n_unsuccess;
  r.num_edge_factors = s->n_edge;
  r.num_plane_factors = s->n_plane;
  r.num_evaluations = s->n_evals;
  r.reserved = 0;
  r.initial_cost = s->initial_cost;
  r.final_cost = s->cost;
  g->report.passes = s->pass + 1;
}

// s: shared-memory copy of the state (everything except the report), g: the HBM original (report sink).
__device__ void lm_advance(LmState* s, LmState* g, const double* sums) {
  const double function_tolerance = 1e-6, gradient_tolerance = 1e-10, parameter_tolerance = 1e-8;
  const double min_relative_decrease = 1e-3, min_radius = 1e-32, max_radius = 1e16;
  const double min_diag = 1e-6, max_diag = 1e32;
  s->n_evals += 1;
  const double new_cost = sums[0];
  if (s->phase == 0) {
    s->n_edge = (int)sums[28];
    s->n_plane = (int)sums[29];
    s->cost = new_cost;
    s->initial_cost = new_cost;
    for (int i = 0; i < 21; ++i) s->H[i] = sums[1 + i];
    for (int i = 0; i < 6; ++i) s->g[i] = sums[22 + i];
    if (s->n_edge + s->n_plane == 0) {  // Ceres: no residual blocks -> parameter blocks dropped -> CONVERGENCE
      lm_terminate(s, g, ILSM_CONVERGENCE);
      return;
    }
    int k = 0;
    for (int a = 0; a < 6; ++a) {
      s->scale[a] = 1.0 / (1.0 + sqrt(s->H[k]));
      k += 6 - a;
    }
  } else {
    double step2 = 0, x2 = 0;
    for (int i = 0; i < 4; ++i) {
      double d = s->xq[i] - s->cq[i];
      step2 += d * d;
      x2 += s->xq[i] * s->xq[i];
    }
    for (int i = 0; i < 3; ++i) {
      double d = s->xt[i] - s->ct[i];
      step2 += d * d;
      x2 += s->xt[i] * s->xt[i];
    }
    if (sqrt(step2) <= parameter_tolerance * (sqrt(x2) + parameter_tolerance)) {
      lm_terminate(s, g, ILSM_CONVERGENCE);
      return;
    }
    double cost_change = s->cost - new_cost;
    if (fabs(cost_change) <= function_tolerance * s->cost) {
      lm_terminate(s, g, ILSM_CONVERGENCE);
      return;
    }
    double rho = cost_change / s->model_cost_change;
    if (rho > min_relative_decrease) {
      for (int i = 0; i < 4; ++i) s->xq[i] = s->cq[i];
      for (int i = 0; i < 3; ++i) s->xt[i] = s->ct[i];
      s->cost = new_cost;
      for (int i = 0; i < 21; ++i) s->H[i] = sums[1 + i];
      for (int i = 0; i < 6; ++i) s->g[i] = sums[22 + i];
      double m = 1.0 - pow(2.0 * rho - 1.0, 3.0);
      s->radius = s->radius / fmax(1.0 / 3.0, m);
      s->radius = fmin(max_radius, s->radius);
      s->decrease_factor = 2.0;
      s->reuse_diag = 0;
      s->n_success += 1;
    } else {
      s->radius = s->radius / s->decrease_factor;
      s->decrease_factor *= 2.0;
      s->reuse_diag = 1;
      s->n_unsuccess += 1;
    }
  }
  // FinalizeIterationAndCheckIfMinimizerCanContinue + ComputeTrustRegionStep, repeated over invalid steps
  for (;;) {
    if (s->iteration >= s->max_iter) {
      lm_terminate(s, g, ILSM_NO_CONVERGENCE);
      return;
    }
    {  // gradient_max_norm = |x - Plus(x, -g)|_inf in the ambient space
      double ng[3] = {-s->g[0], -s->g[1], -s->g[2]}, qp[4];
      quat_plus_d(s->xq, ng, qp);
      double m = 0;
      for (int i = 0; i < 4; ++i) m = fmax(m, fabs(s->xq[i] - qp[i]));
      for (int i = 0; i < 3; ++i) m = fmax(m, fabs(s->g[3 + i]));
      if (m <= gradient_tolerance) {
        lm_terminate(s, g, ILSM_CONVERGENCE);
        return;
      }
    }
    if (s->radius <= min_radius) {
      lm_terminate(s, g, ILSM_CONVERGENCE);
      return;
    }
    s->iteration += 1;
    double Hs[6][6], gs[6], A[6][6], y[6], step[6];
    {
      int k = 0;
      for (int a = 0; a < 6; ++a)
        for (int b = a; b < 6; ++b) {
          double v = s->H[k++] * s->scale[a] * s->scale[b];
          Hs[a][b] = v;
          Hs[b][a] = v;
        }
      for (int a = 0; a < 6; ++a) gs[a] = s->g[a] * s->scale[a];
    }
    if (!s->reuse_diag)
      for (int a = 0; a < 6; ++a) s->diag[a] = fmin(fmax(Hs[a][a], min_diag), max_diag);
    for (int a = 0; a < 6; ++a)
      for (int b = 0; b < 6; ++b) A[a][b] = Hs[a][b] + (a == b ? s->diag[a] / s->radius : 0.0);
    bool ok = chol_solve6_d(A, gs, y);
    s->reuse_diag = 1;
    double mcc = 0;
    if (ok) {
      double sg = 0, sHs = 0;
      for (int a = 0; a < 6; ++a) step[a] = -y[a];
      for (int a = 0; a < 6; ++a) {
        sg += step[a] * gs[a];
        double hv = 0;
        for (int b = 0; b < 6; ++b) hv += Hs[a][b] * step[b];
        sHs += step[a] * hv;
        ok = ok && isfinite(step[a]);
      }
      mcc = -(sg + 0.5 * sHs);
    }
    if (!ok || !(mcc > 0.0)) {  // HandleInvalidStep
      s->n_unsuccess += 1;
      if (++s->invalid_run >= 5) {
        lm_terminate(s, g, ILSM_FAILURE);
        return;
      }
      s->radius = s->radius / s->decrease_factor;
      s->decrease_factor *= 2.0;
      s->reuse_diag = 1;
      continue;
    }
    s->invalid_run = 0;
    s->model_cost_change = mcc;
    double delta[6];
    for (int a = 0; a < 6; ++a) delta[a] = step[a] * s->scale[a];
    quat_plus_d(s->xq, delta, s->cq);
    for (int i = 0; i < 3; ++i) s->ct[i] = s->xt[i] + delta[3 + i];
    s->phase = 1;
    return;
  }
}

// ---------------------------------------------------------------------------------------------------
// evaluation kernel
// ---------------------------------------------------------------------------------------------------
__device__ __forceinline__ void acc_row(double (&acc)[kNumSums], const double (&j)[6], double r) {
  int k = 1;
#pragma unroll
  for (int a = 0; a < 6; ++a) {
#pragma unroll
    for (int b = a; b < 6; ++b) acc[k++] += j[a] * j[b];
    acc[22 + a] += j[a] * r;
  }
}

// mode 0: advance the LM state machine; mode 1: evaluate only (sums -> eval_out)
__global__ void __launch_bounds__(256) eval_kernel(FactorView fv, int n, LmState* st, double* __restrict__ partials,
                                                   double* __restrict__ eval_out, int mode) {
  if (mode == 0 && st->status != 0) return;
  __shared__ double red[8][kSumStride];
  __shared__ double tot[kSumStride];
  __shared__ int is_last;
  constexpr int kCoreWords = (int)(offsetof(LmState, report) / 8);
  __shared__ double core[kCoreWords];

  double acc[kNumSums];
#pragma unroll
  for (int i = 0; i < kNumSums; ++i) acc[i] = 0.0;

  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  const int type = i < n ? fv.type[i] : 0;
  if (type != 0) {
    const double huber_a = st->huber_a;
    double q[4] = {st->cq[0], st->cq[1], st->cq[2], st->cq[3]};
    const float4 pf = fv.p[i];
    const double4 fa = fv.a[i];
    D3 Rp = quat_rotate(q, d3((double)pf.x, (double)pf.y, (double)pf.z));
    D3 lp = d3(Rp.x + st->ct[0], Rp.y + st->ct[1], Rp.z + st->ct[2]);
    double r[3] = {0, 0, 0}, J[3][6];
    int nres;
    if (type == 1) {
      const double4 fb = fv.b[i];
      D3 u = d3(lp.x - fa.x, lp.y - fa.y, lp.z - fa.z), v = d3(lp.x - fb.x, lp.y - fb.y, lp.z - fb.z);
      D3 nu = d3(u.y * v.z - u.z * v.y, u.z * v.x - u.x * v.z, u.x * v.y - u.y * v.x);
      double dex = fa.x - fb.x, dey = fa.y - fb.y, dez = fa.z - fb.z;
      double dn = sqrt(dex * dex + dey * dey + dez * dez);
      r[0] = nu.x / dn, r[1] = nu.y / dn, r[2] = nu.z / dn;
      double mx = -dex / dn, my = -dey / dn, mz = -dez / dn;  // (b - a)/|a-b|
      double Jt[3][3] = {{0, -mz, my}, {mz, 0, -mx}, {-my, mx, 0}};
      double S[3][3] = {{0, 2 * Rp.z, -2 * Rp.y}, {-2 * Rp.z, 0, 2 * Rp.x}, {2 * Rp.y, -2 * Rp.x, 0}};
#pragma unroll
      for (int a = 0; a < 3; ++a)
#pragma unroll
        for (int c = 0; c < 3; ++c) {
          J[a][c] = Jt[a][0] * S[0][c] + Jt[a][1] * S[1][c] + Jt[a][2] * S[2][c];
          J[a][3 + c] = Jt[a][c];
        }
      nres = 3;
      acc[28] = 1.0;
    } else {
      r[0] = fa.x * lp.x + fa.y * lp.y + fa.z * lp.z + fa.w;
      J[0][0] = 2.0 * (Rp.y * fa.z - Rp.z * fa.y);
      J[0][1] = 2.0 * (Rp.z * fa.x - Rp.x * fa.z);
      J[0][2] = 2.0 * (Rp.x * fa.y - Rp.y * fa.x);
      J[0][3] = fa.x, J[0][4] = fa.y, J[0][5] = fa.z;
#pragma unroll
      for (int a = 1; a < 3; ++a)
#pragma unroll
        for (int c = 0; c < 6; ++c) J[a][c] = 0.0;
      nres = 1;
      acc[29] = 1.0;
    }
    double sq = r[0] * r[0] + r[1] * r[1] + r[2] * r[2];
    double rho0 = sq, rho1 = 1.0;
    if (huber_a > 0.0 && sq > huber_a * huber_a) {  // ceres::HuberLoss + Corrector (rho'' <= 0)
      double rr = sqrt(sq);
      rho0 = 2.0 * huber_a * rr - huber_a * huber_a;
      rho1 = fmax(2.2250738585072014e-308, huber_a / rr);
    }
    const double sc = sqrt(rho1);
    acc[0] = 0.5 * rho0;
#pragma unroll
    for (int a = 0; a < 3; ++a) {
      if (a < nres) {
        double jr[6];
#pragma unroll
        for (int c = 0; c < 6; ++c) jr[c] = sc * J[a][c];
        acc_row(acc, jr, sc * r[a]);
      }
    }
  }
  // warp shuffle reduction, then cross-warp through shared memory (fixed order => deterministic)
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
  for (int k = 0; k < kNumSums; ++k) {
    double v = acc[k];
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) v += __shfl_xor_sync(0xffffffffu, v, off);
    if (lane == 0) red[warp][k] = v;
  }
  __syncthreads();
  const int nwarps = blockDim.x >> 5;
  if (threadIdx.x < kNumSums) {
    double v = 0;
    for (int w = 0; w < nwarps; ++w) v += red[w][threadIdx.x];
    partials[(size_t)blockIdx.x * kSumStride + threadIdx.x] = v;
  }
  __threadfence();
  __syncthreads();
  if (threadIdx.x == 0) {
    unsigned t = atomicAdd(&st->ticket, 1u);
    is_last = (t == gridDim.x - 1);
  }
  __syncthreads();
  if (!is_last) return;
  __threadfence();
  if (threadIdx.x < kNumSums) {
    double v = 0;
    for (unsigned b = 0; b < gridDim.x; ++b) v += __ldcg(partials + (size_t)b * kSumStride + threadIdx.x);
    tot[threadIdx.x] = v;
  }
  if (mode == 1) {
    __syncthreads();
    if (threadIdx.x < kNumSums) eval_out[threadIdx.x] = tot[threadIdx.x];
    if (threadIdx.x == 0) st->ticket = 0u;
    return;
  }
  {
    const double* src = reinterpret_cast<const double*>(st);
    for (int w = threadIdx.x; w < kCoreWords; w += blockDim.x) core[w] = src[w];
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    LmState* s = reinterpret_cast<LmState*>(core);  // only the fields before `report` are touched through s
    s->ticket = 0u;
    lm_advance(s, st, tot);
  }
  __syncthreads();
  {
    double* dst = reinterpret_cast<double*>(st);
    for (int w = threadIdx.x; w < kCoreWords; w += blockDim.x) dst[w] = core[w];
  }
}

// ---------------------------------------------------------------------------------------------------
// host side
// ---------------------------------------------------------------------------------------------------
static FactorView factor_view(FactorBufs& f, bool want_knn) {
  FactorView v;
  v.type = f.type.p;
  v.p = f.p.p;
  v.a = f.a.p;
  v.b = f.b.p;
  v.knn_idx = want_knn ? f.knn_idx.p : nullptr;
  v.knn_d2 = want_knn ? f.knn_d2.p : nullptr;
  return v;
}

int Ctx::associate_dev(Map* mc, Map* ms, const float* d_corner, int nc, const float* d_surf, int ns, int stride_bytes,
                       const ilsm_reg_opts& o, bool begin_solve, int pass, bool want_knn) {
  const int n = nc + ns;
  int rc;
  if ((rc = fac.type.reserve(n + 1)) || (rc = fac.p.reserve(n + 1)) || (rc = fac.a.reserve(n + 1)) ||
      (rc = fac.b.reserve(n + 1)))
    return rc;
  if (want_knn && ((rc = fac.knn_idx.reserve((size_t)n * 5 + 1)) || (rc = fac.knn_d2.reserve((size_t)n * 5 + 1))))
    return rc;
  fac.n = n;
  fac.nc = nc;
  AssocParams prm;
  prm.gate_sq = o.knn_gate_sq;
  prm.line_ratio = o.line_eig_ratio;
  prm.plane_tol = o.plane_tol;
  prm.begin_solve = begin_solve ? 1 : 0;
  prm.pass = pass;
  prm.max_iter = o.max_num_iterations;
  prm.huber_a = o.huber_a;
  GridView gc = mc->view(), gs = ms->view();
  FactorView fv = factor_view(fac, want_knn);
  const int T = 256;
  const int stride_f = stride_bytes / 4;
  if (n == 0) {
    // still (re)arm the LM state so that the following evaluation terminates with "no residuals"
    associate_kernel<32><<<1, 32, 0, stream>>>(gc, gs, d_corner, 0, d_surf, 0, stride_f, lm.p, prm, fv);
  } else if ((long long)n * 32 <= (long long)sm_count * 2048 * 2) {
    associate_kernel<32><<<(n + T / 32 - 1) / (T / 32), T, 0, stream>>>(gc, gs, d_corner, nc, d_surf, ns, stride_f, lm.p,
                                                                         prm, fv);
  } else {
    associate_kernel<8><<<(n + T / 8 - 1) / (T / 8), T, 0, stream>>>(gc, gs, d_corner, nc, d_surf, ns, stride_f, lm.p,
                                                                       prm, fv);
  }
  count_launches(1);
  return check_launch("associate");
}

int Ctx::eval_launch(int count) {
  const int T = 256;
  int blocks = (fac.n + T - 1) / T;
  if (blocks < 1) blocks = 1;
  int rc;
  if ((rc = partials.reserve((size_t)blocks * kSumStride + kSumStride))) return rc;
  FactorView fv = factor_view(fac, false);
  for (int k = 0; k < count; ++k)
    eval_kernel<<<blocks, T, 0, stream>>>(fv, fac.n, lm.p, partials.p, nullptr, 0);
  count_launches(count);
  return check_launch("eval");
}

int Ctx::register_dev(Map* mc, Map* ms, const float* d_corner, int nc, const float* d_surf, int ns, int stride_bytes,
                      const ilsm_reg_opts& o) {
  for (int pass = 0; pass < o.outer_iterations; ++pass) {
    int rc = associate_dev(mc, ms, d_corner, nc, d_surf, ns, stride_bytes, o, true, pass, false);
    if (rc) return rc;
    if ((rc = eval_launch(1 + o.max_num_iterations))) return rc;
  }
  return ILSM_OK;
}

// stand-alone evaluation (ilsm_eval_normal_eq): candidate pose must already be in lm->cq/ct, huber in lm->huber_a
int eval_only_launch(Ctx* c, double* d_out) {
  const int T = 256;
  int blocks = (c->fac.n + T - 1) / T;
  if (blocks < 1) blocks = 1;
  int rc;
  if ((rc = c->partials.reserve((size_t)blocks * kSumStride + kSumStride))) return rc;
  FactorView fv = factor_view(c->fac, false);
  eval_kernel<<<blocks, T, 0, c->stream>>>(fv, c->fac.n, c->lm.p, c->partials.p, d_out, 1);
  count_launches(1);
  return check_launch("eval_only");
}

// copy factor SoA -> AoS records on the device for ilsm_associate's host output
__global__ void factors_export_kernel(FactorView fv, int n, int nc, ilsm_factor* out) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  ilsm_factor f;
  f.type = fv.type[i];
  f.src = i < nc ? i : i - nc;
  float4 p = fv.p[i];
  double4 a = fv.a[i], b = fv.b[i];
  f.p[0] = p.x, f.p[1] = p.y, f.p[2] = p.z;
  f.a[0] = a.x, f.a[1] = a.y, f.a[2] = a.z;
  if (f.type == 2) {
    f.b[0] = a.w, f.b[1] = 0, f.b[2] = 0;
  } else {
    f.b[0] = b.x, f.b[1] = b.y, f.b[2] = b.z;
  }
  out[i] = f;
}

int factors_export(Ctx* c, ilsm_factor* d_out) {
  if (c->fac.n == 0) return ILSM_OK;
  FactorView fv = factor_view(c->fac, false);
  factors_export_kernel<<<(c->fac.n + 255) / 256, 256, 0, c->stream>>>(fv, c->fac.n, c->fac.nc, d_out);
  count_launches(1);
  return check_launch("factors_export");
}

}  // namespace ilsm
