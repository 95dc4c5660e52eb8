// ref_laserodom.cpp -- TEST INFRASTRUCTURE.  Runs the reference's own scan-to-scan data association
// (src/laserOdometry.cpp:417-713 with TransformToStart :147-172, cut out of the node's spin loop by
// oracle/patches/laserodom_extract.py) and returns the residual blocks it built: for every sharp point the two previous-frame
// points of its edge factor, for every flat point the three of its plane factor.  The oracle's restatement of the closest-
// point search, the +-2.5-ring walks, the distance gate and TransformToStart is checked against them.
// Stand-ins (not the reference): pcl::PointCloud, pcl::KdTreeFLANN (exact brute-force 1-NN, float L2 as FLANN's L2_Simple
// evaluates it, lowest index on ties), Eigen (oracle/shims/eigen3: Eigen 3.3's formulas), ceres::Problem (records the
// blocks) and ceres::Solve (does nothing: both passes of the loop associate at the pose the caller set).
// Built only into oracle/_ref/libref_laserodom.so (git-ignored); nothing in the product path links it.
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstring>
#include <memory>
#include <string>
#include <vector>

#include "tic_toc.h"
#include "lidarFeaturePointsFunction.hpp"  // the reference's functors (LidarEdgeFactor / LidarPlaneFactor), on the shims

void (*ceres::Problem::sink)(ceres::CostFunction*, void*) = nullptr;
void* ceres::Problem::sink_arg = nullptr;

namespace pcl {
template <typename PointT>
struct PointCloud {
  typedef std::shared_ptr<PointCloud<PointT>> Ptr;
  std::vector<PointT> points;
  size_t size() const { return points.size(); }
};
template <typename PointT>
struct KdTreeFLANN {
  typedef std::shared_ptr<KdTreeFLANN<PointT>> Ptr;
  typename PointCloud<PointT>::Ptr cloud;
  void setInputCloud(const typename PointCloud<PointT>::Ptr& c) { cloud = c; }
  int nearestKSearch(const PointT& q, int k, std::vector<int>& idx, std::vector<float>& d2) const {
    idx.assign(k, 0), d2.assign(k, 0.f);  // k == 1 in laserOdometry.cpp
    float best = INFINITY;
    int bi = -1;
    for (int i = 0; i < (int)cloud->points.size(); ++i) {
      const PointT& p = cloud->points[i];
      const float dx = q.x - p.x, dy = q.y - p.y, dz = q.z - p.z;
      const float d = (dx * dx + dy * dy) + dz * dz;
      if (d < best) best = d, bi = i;
    }
    if (bi < 0) {
      d2[0] = INFINITY;
      return 0;
    }
    idx[0] = bi, d2[0] = best;
    return 1;
  }
};
}  // namespace pcl
typedef pcl::PointXYZI PointType;  // parameters.h_ouster

#include "globals.inc"
#include "transform.inc"

struct Captured {
  std::vector<double> edge, plane;  // edge: curr 3, a 3, b 3, s;  plane: curr 3, j 3, l 3, m 3, s
};
static void capture(ceres::CostFunction* f, void* arg) {
  Captured& c = *static_cast<Captured*>(arg);
  if (auto* e = dynamic_cast<ceres::AutoDiffCostFunction<LidarEdgeFactor, 3, 4, 3>*>(f)) {
    const LidarEdgeFactor& k = *e->functor_;
    for (const Eigen::Vector3d* v : {&k.curr_point, &k.last_point_a, &k.last_point_b}) c.edge.insert(c.edge.end(), {v->x(), v->y(), v->z()});
    c.edge.push_back(k.s);
  } else if (auto* p = dynamic_cast<ceres::AutoDiffCostFunction<LidarPlaneFactor, 1, 4, 3>*>(f)) {
    const LidarPlaneFactor& k = *p->functor_;
    for (const Eigen::Vector3d* v : {&k.curr_point, &k.last_point_j, &k.last_point_l, &k.last_point_m})
      c.plane.insert(c.plane.end(), {v->x(), v->y(), v->z()});
    c.plane.push_back(k.s);
  }
}

static void fill(pcl::PointCloud<PointType>& c, const float* xyzi, int n) {
  c.points.resize(n);
  for (int i = 0; i < n; ++i) c.points[i].x = xyzi[4 * i], c.points[i].y = xyzi[4 * i + 1], c.points[i].z = xyzi[4 * i + 2], c.points[i].intensity = xyzi[4 * i + 3];
}

// One run of the reference's association loop at pose (q = x, y, z, w; t).  All clouds packed xyzi (intensity = scan id +
// relTime, as scanRegistration emits them).  edge_out: n_sharp x 10 doubles max, plane_out: n_flat x 13; counts[0..1] =
// blocks of the LAST pass (both passes see the same pose: the stand-in Solve does not move it), counts[2..3] = the
// reference's own corner_correspondence / plane_correspondence counters.
extern "C" int ref_laserodom_associate(const float* last_corner, int n_lc, const float* last_surf, int n_ls, const float* sharp, int n_sharp,
                                       const float* flat, int n_flat, const double* q_xyzw, const double* t_xyz, double* edge_out,
                                       double* plane_out, int32_t* counts) {
  fill(*laserCloudCornerLast, last_corner, n_lc), fill(*laserCloudSurfLast, last_surf, n_ls);
  fill(*cornerPointsSharp, sharp, n_sharp), fill(*surfPointsFlat, flat, n_flat);
  kdtreeCornerLast->setInputCloud(laserCloudCornerLast), kdtreeSurfLast->setInputCloud(laserCloudSurfLast);
  for (int i = 0; i < 4; ++i) para_q[i] = q_xyzw[i];
  for (int i = 0; i < 3; ++i) para_t[i] = t_xyz[i];
  int cornerPointsSharpNum = cornerPointsSharp->points.size();  // laserOdometry.cpp:394,397
  int surfPointsFlatNum = surfPointsFlat->points.size();
  bool use_aloam = true;
  Captured cap;  // both passes of the loop append here; the stand-in Solve does not move the pose, so they are identical
  ceres::Problem::sink_arg = &cap;
  ceres::Problem::sink = capture;
#define printf(...) ((void)0)  // the body reports its solver time on stdout every pass
  {
#include "body.inc"
  }
#undef printf
  ceres::Problem::sink = nullptr;
  const size_t ne = cap.edge.size() / 10 / 2, np = cap.plane.size() / 13 / 2;
  if (cap.edge.size() != ne * 20 || cap.plane.size() != np * 26) return -1;
  if (memcmp(cap.edge.data(), cap.edge.data() + ne * 10, ne * 10 * sizeof(double)) ||
      memcmp(cap.plane.data(), cap.plane.data() + np * 13, np * 13 * sizeof(double)))
    return -2;
  memcpy(edge_out, cap.edge.data(), ne * 10 * sizeof(double));
  memcpy(plane_out, cap.plane.data(), np * 13 * sizeof(double));
  counts[0] = (int)ne, counts[1] = (int)np, counts[2] = corner_correspondence, counts[3] = plane_correspondence;
  return 0;
}
