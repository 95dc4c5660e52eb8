// ref_lasermapping.cpp -- TEST INFRASTRUCTURE.  Runs the reference's own map maintenance (src/laserMapping.cpp: the rolling
// 21x21x11 cube window :327-565, the 5x5x3 gather :566-593, the stack VoxelGrids :595-603, transformUpdate :875, the insertion
// :878-945 and the per-cube VoxelGrid of the valid cubes :984-1002, cut out of process() by
// oracle/patches/lasermapping_extract.py) frame by frame, so that the oracle's restatement of the window arithmetic -- centre
// index with its negative fix-up, the six roll directions, the valid-cube order, the cube index of every inserted point --
// can be compared with the reference code cube by cube.  The guarded optimisation (:624-873) is left out: both sides of the
// comparison keep the pose transformAssociateToMap predicts.
// Stand-ins (not the reference): pcl::PointCloud, pcl::VoxelGrid (the oracle's restatement, orc_voxelgrid), Eigen
// (oracle/shims/eigen3, the oracle's quaternion arithmetic), pcl::KdTreeFLANN (declared by the globals, unused here).
// Built only into oracle/_ref/libref_lasermapping.so (git-ignored); nothing in the product path links it.
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstring>
#include <memory>
#include <vector>

#include <eigen3/Eigen/Dense>
#include <pcl/point_types.h>

#include "tic_toc.h"

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
struct KdTreeFLANN {
  typedef std::shared_ptr<KdTreeFLANN<PointT>> Ptr;
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

extern "C" int ref_lasermapping_cube(int which, int index, float* out_xyzi, int cap) {
  const pcl::PointCloud<PointType>& c = which == 0 ? *laserCloudCornerArray[index] : *laserCloudSurfArray[index];
  const int n = (int)c.points.size();
  for (int i = 0; i < n && i < cap; ++i) out_xyzi[4 * i] = c.points[i].x, out_xyzi[4 * i + 1] = c.points[i].y, out_xyzi[4 * i + 2] = c.points[i].z, out_xyzi[4 * i + 3] = c.points[i].intensity;
  return n;
}
