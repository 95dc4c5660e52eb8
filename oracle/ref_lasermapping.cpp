// ref_lasermapping.cpp -- TEST INFRASTRUCTURE.  Runs the reference's own map maintenance (src/laserMapping.cpp: the rolling
// 21x21x11 cube window :327-565, the 5x5x3 gather :566-593, the stack VoxelGrids :595-603, transformUpdate :875, the insertion
// :878-945 and the per-cube VoxelGrid of the valid cubes :984-1002, cut out of process() by
// oracle/patches/lasermapping_extract.py) frame by frame, so that the oracle's restatement of the window arithmetic -- centre
// index with its negative fix-up, the six roll directions, the valid-cube order, the cube index of every inserted point --
// can be compared with the reference code cube by cube.  The guarded optimisation (:624-873) is left out: both sides of the
// comparison keep the pose transformAssociateToMap predicts.
// ref_lasermapping_associate runs that block on its own with a recording ceres::Problem: the 5-NN gate, the line test
// (lambda_2 > 3 lambda_1, point_a / point_b = centre +- 0.1 v_2) and the plane test (all five within 0.2) of :640-796.
// Stand-ins (not the reference): pcl::PointCloud, pcl::VoxelGrid (the oracle's restatement, orc_voxelgrid), pcl::KdTreeFLANN
// (exact brute force), Eigen (oracle/shims/eigen3: the oracle's quaternion arithmetic, and its eigen / QR kernels behind
// SelfAdjointEigenSolver / colPivHouseholderQr), ceres::Problem (records) and ceres::Solve (does nothing).
// Built only into oracle/_ref/libref_lasermapping.so (git-ignored); nothing in the product path links it.
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstring>
#include <memory>
#include <vector>

#include "tic_toc.h"
#include "lidarFeaturePointsFunction.hpp"  // the reference's functors (LidarEdgeFactor / LidarPlaneNormFactor), on the shims

void (*ceres::Problem::sink)(ceres::CostFunction*, void*) = nullptr;
void* ceres::Problem::sink_arg = nullptr;

extern "C" int orc_voxelgrid(const float* in, int n, int stride_bytes, int ioff, float leaf, float* out_xyzi);  // ilsm_oracle_frontend.cpp

namespace pcl {
template <typename PointT>
struct PointCloud {
  typedef std::shared_ptr<PointCloud<PointT>> Ptr;
  std::vector<PointT> points;
  size_t size() const { return points.size(); }
  void clear() { points.clear(); }
  void push_back(const PointT& p) { points.push_back(p); }
  PointCloud& operator+=(const PointCloud& o) {
    points.insert(points.end(), o.points.begin(), o.points.end());
    return *this;
  }
};
template <typename PointT>
struct KdTreeFLANN {  // exact brute-force k-NN, float L2 as FLANN's L2_Simple evaluates it, ascending (distance, index)
  typedef std::shared_ptr<KdTreeFLANN<PointT>> Ptr;
  typename PointCloud<PointT>::Ptr cloud;
  void setInputCloud(const typename PointCloud<PointT>::Ptr& c) { cloud = c; }
  int nearestKSearch(const PointT& q, int k, std::vector<int>& idx, std::vector<float>& d2) const {
    idx.assign(k, 0), d2.assign(k, INFINITY);
    int found = 0;
    for (int i = 0; i < (int)cloud->points.size(); ++i) {
      const PointT& p = cloud->points[i];
      const float dx = q.x - p.x, dy = q.y - p.y, dz = q.z - p.z;
      const float d = (dx * dx + dy * dy) + dz * dz;
      if (found < k || d < d2[k - 1]) {  // insertion keeps ties in index order
        int j = found < k ? found : k - 1;
        while (j > 0 && d2[j - 1] > d) d2[j] = d2[j - 1], idx[j] = idx[j - 1], --j;
        d2[j] = d, idx[j] = i;
        if (found < k) ++found;
      }
    }
    return found;
  }
};
template <typename PointT>
struct VoxelGrid {  // the ORACLE's restatement of PCL's filter, not PCL
  typename PointCloud<PointT>::Ptr in;
  float leaf = 0.f;
  void setInputCloud(const typename PointCloud<PointT>::Ptr& c) { in = c; }
  void setLeafSize(float lx, float, float) { leaf = lx; }
  void filter(PointCloud<PointT>& out) {
    const int n = (int)in->points.size();
    std::vector<float> a((size_t)n * 4 + 4), b((size_t)n * 4 + 4);
    for (int i = 0; i < n; ++i) a[4 * i] = in->points[i].x, a[4 * i + 1] = in->points[i].y, a[4 * i + 2] = in->points[i].z, a[4 * i + 3] = in->points[i].intensity;
    const int m = n ? orc_voxelgrid(a.data(), n, 16, 3, leaf, b.data()) : 0;
    out.points.resize(m);
    for (int i = 0; i < m; ++i) out.points[i].x = b[4 * i], out.points[i].y = b[4 * i + 1], out.points[i].z = b[4 * i + 2], out.points[i].intensity = b[4 * i + 3];
  }
};
}  // namespace pcl
typedef pcl::PointXYZI PointType;  // parameters.h_ouster
#define ROS_WARN(...) ((void)0)

#include "globals.inc"
#include "transform.inc"

static void fill(pcl::PointCloud<PointType>& c, const float* xyzi, int n) {
  c.points.resize(n);
  for (int i = 0; i < n; ++i) c.points[i].x = xyzi[4 * i], c.points[i].y = xyzi[4 * i + 1], c.points[i].z = xyzi[4 * i + 2], c.points[i].intensity = xyzi[4 * i + 3];
}

// main() :1140-1160: leaf sizes of the two filters, one empty cloud per cube; state back to its initial values
extern "C" void ref_lasermapping_reset(float line_res, float plane_res) {
  downSizeFilterCorner.setLeafSize(line_res, line_res, line_res);
  downSizeFilterSurf.setLeafSize(plane_res, plane_res, plane_res);
  for (int i = 0; i < laserCloudNum; i++) {
    laserCloudCornerArray[i].reset(new pcl::PointCloud<PointType>());
    laserCloudSurfArray[i].reset(new pcl::PointCloud<PointType>());
  }
  laserCloudCenWidth = 10, laserCloudCenHeight = 10, laserCloudCenDepth = 5;
  const double p0[7] = {0, 0, 0, 1, 0, 0, 0};
  memcpy(parameters, p0, sizeof(p0));
  q_wmap_wodom = Eigen::Quaterniond(1, 0, 0, 0), t_wmap_wodom = Eigen::Vector3d(0, 0, 0);
}

// One process() iteration for the feature clouds of a frame (packed xyzi) and its odometry pose (q = x, y, z, w; t).
// out: qt_w[7] the pose the frame was inserted with, cen[3] the window centre indices, n_valid + valid[125] the cube
// indices of the 5x5x3 neighbourhood in the reference's order, sizes[4] = map corner / map surf / stack corner / stack surf.
extern "C" void ref_lasermapping_frame(const float* corner_last, int nc, const float* surf_last, int ns, const double* q_odom_xyzw,
                                       const double* t_odom, double* qt_w, int32_t* cen, int32_t* n_valid, int32_t* valid,
                                       int32_t* sizes) {
  fill(*laserCloudCornerLast, corner_last, nc), fill(*laserCloudSurfLast, surf_last, ns);
  q_wodom_curr = Eigen::Quaterniond(q_odom_xyzw[3], q_odom_xyzw[0], q_odom_xyzw[1], q_odom_xyzw[2]);  // :303-309
  t_wodom_curr = Eigen::Vector3d(t_odom[0], t_odom[1], t_odom[2]);
#define printf(...) ((void)0)
  {
#include "body.inc"
    *n_valid = laserCloudValidNum;
    for (int i = 0; i < laserCloudValidNum; i++) valid[i] = laserCloudValidInd[i];
    sizes[0] = laserCloudCornerFromMapNum, sizes[1] = laserCloudSurfFromMapNum;
    sizes[2] = laserCloudCornerStackNum, sizes[3] = laserCloudSurfStackNum;
  }
#undef printf
  for (int i = 0; i < 7; ++i) qt_w[i] = parameters[i];
  cen[0] = laserCloudCenWidth, cen[1] = laserCloudCenHeight, cen[2] = laserCloudCenDepth;
}

// The guarded optimisation block (:624-873) on its own, at pose qt (q = x, y, z, w; t) against the given map and stack clouds
// (packed xyzi): returns the residual blocks of the last pass -- edge: curr 3, point_a 3, point_b 3, s; plane: curr 3, unit
// normal 3, negative_OA_dot_norm.  The stand-in Solve leaves the pose alone, so both passes build the same blocks (checked).
extern "C" int ref_lasermapping_associate(const float* map_corner, int n_mc, const float* map_surf, int n_ms, const float* stack_corner,
                                          int n_sc, const float* stack_surf, int n_ss, const double* qt, double* edge_out, double* plane_out,
                                          int32_t* counts) {
  fill(*laserCloudCornerFromMap, map_corner, n_mc), fill(*laserCloudSurfFromMap, map_surf, n_ms);
  pcl::PointCloud<PointType>::Ptr laserCloudCornerStack(new pcl::PointCloud<PointType>());  // locals of process() (:595-603)
  pcl::PointCloud<PointType>::Ptr laserCloudSurfStack(new pcl::PointCloud<PointType>());
  fill(*laserCloudCornerStack, stack_corner, n_sc), fill(*laserCloudSurfStack, stack_surf, n_ss);
  int laserCloudCornerFromMapNum = n_mc, laserCloudSurfFromMapNum = n_ms;
  int laserCloudCornerStackNum = n_sc, laserCloudSurfStackNum = n_ss;
  for (int i = 0; i < 7; ++i) parameters[i] = qt[i];
  struct Captured {
    std::vector<double> edge, plane;
  } cap;
  ceres::Problem::sink_arg = &cap;
  ceres::Problem::sink = [](ceres::CostFunction* f, void* arg) {
    Captured& c = *static_cast<Captured*>(arg);
    if (auto* e = dynamic_cast<ceres::AutoDiffCostFunction<LidarEdgeFactor, 3, 4, 3>*>(f)) {
      const LidarEdgeFactor& k = *e->functor_;
      for (const Eigen::Vector3d* v : {&k.curr_point, &k.last_point_a, &k.last_point_b}) c.edge.insert(c.edge.end(), {v->x(), v->y(), v->z()});
      c.edge.push_back(k.s);
    } else if (auto* p = dynamic_cast<ceres::AutoDiffCostFunction<LidarPlaneNormFactor, 1, 4, 3>*>(f)) {
      const LidarPlaneNormFactor& k = *p->functor_;
      for (const Eigen::Vector3d* v : {&k.curr_point, &k.plane_unit_norm}) c.plane.insert(c.plane.end(), {v->x(), v->y(), v->z()});
      c.plane.push_back(k.negative_OA_dot_norm);
    }
  };
#define printf(...) ((void)0)
  {
#include "optimise.inc"
  }
#undef printf
  ceres::Problem::sink = nullptr;
  const size_t ne = cap.edge.size() / 10 / 2, np = cap.plane.size() / 7 / 2;
  if (cap.edge.size() != ne * 20 || cap.plane.size() != np * 14) return -1;
  if (memcmp(cap.edge.data(), cap.edge.data() + ne * 10, ne * 10 * sizeof(double)) ||
      memcmp(cap.plane.data(), cap.plane.data() + np * 7, np * 7 * sizeof(double)))
    return -2;
  memcpy(edge_out, cap.edge.data(), ne * 10 * sizeof(double));
  memcpy(plane_out, cap.plane.data(), np * 7 * sizeof(double));
  counts[0] = (int)ne, counts[1] = (int)np;
  return 0;
}

extern "C" int ref_lasermapping_cube(int which, int index, float* out_xyzi, int cap) {
  const pcl::PointCloud<PointType>& c = which == 0 ? *laserCloudCornerArray[index] : *laserCloudSurfArray[index];
  const int n = (int)c.points.size();
  for (int i = 0; i < n && i < cap; ++i) out_xyzi[4 * i] = c.points[i].x, out_xyzi[4 * i + 1] = c.points[i].y, out_xyzi[4 * i + 2] = c.points[i].z, out_xyzi[4 * i + 3] = c.points[i].intensity;
  return n;
}
