// ref_scanreg.cpp -- TEST INFRASTRUCTURE.  Runs the reference's own LOAM front end (src/scanRegistration.cpp:227-589,
// cut out of its ROS node by oracle/patches/scanreg_extract.py) on a caller-supplied cloud, so that the oracle's
// restatement of removeClosedPointCloud, the 64-ring / relTime tagging, the curvature, the per-segment std::sort and the
// sharp / less-sharp / flat / less-flat picking can be checked against the reference code itself.
// What is NOT the reference here: pcl::PointCloud (a std::vector stand-in with PCL's interface as far as the body uses
// it) and pcl::VoxelGrid, which forwards to the oracle's restatement (orc_voxelgrid) -- PCL is not installed, so the
// per-ring VoxelGrid(0.2) of :580-589 stays unpinned and the less-flat output is compared BEFORE that filter as well.
// Built only into oracle/_ref/libref_scanreg.so (git-ignored); nothing in the product path links it.
#include <algorithm>
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <memory>
#include <vector>

#include "tic_toc.h"  // the reference's own header (src/tic_toc.h: std::chrono only)

extern "C" int orc_voxelgrid(const float* in, int n, int stride_bytes, int ioff, float leaf, float* out_xyzi);  // ilsm_oracle_frontend.cpp

namespace pcl {
struct alignas(16) PointXYZI {  // PCL's layout: 32 bytes
  float x = 0.f, y = 0.f, z = 0.f, data3 = 1.f;
  float intensity = 0.f, pad[3] = {0.f, 0.f, 0.f};
};
struct PCLHeader {
  uint32_t seq = 0;
  uint64_t stamp = 0;
};
template <typename PointT>
struct PointCloud {
  typedef std::shared_ptr<PointCloud<PointT>> Ptr;
  typedef std::shared_ptr<const PointCloud<PointT>> ConstPtr;
  PCLHeader header;
  std::vector<PointT> points;
  uint32_t width = 0, height = 0;
  bool is_dense = true;
  size_t size() const { return points.size(); }
  void push_back(const PointT& p) {
    points.push_back(p);
    width = (uint32_t)points.size(), height = 1;
  }
  PointCloud& operator+=(const PointCloud& o) {
    points.insert(points.end(), o.points.begin(), o.points.end());
    width = (uint32_t)points.size(), height = 1;
    return *this;
  }
};
// the per-ring VoxelGrid(0.2) of scanRegistration.cpp:580-589: the ORACLE's restatement of PCL's filter, not PCL
static std::vector<float> g_lessflat_raw;  // the less-flat points as collected, before any VoxelGrid (rings concatenated)
template <typename PointT>
struct VoxelGrid {
  typename PointCloud<PointT>::Ptr in;
  float leaf = 0.f;
  void setInputCloud(const typename PointCloud<PointT>::Ptr& c) { in = c; }
  void setLeafSize(float lx, float, float) { leaf = lx; }
  void filter(PointCloud<PointT>& out) {
    const int n = (int)in->points.size();
    std::vector<float> a((size_t)n * 4 + 4), b((size_t)n * 4 + 4);
    for (int i = 0; i < n; ++i) {
      a[4 * i] = in->points[i].x, a[4 * i + 1] = in->points[i].y, a[4 * i + 2] = in->points[i].z, a[4 * i + 3] = in->points[i].intensity;
      g_lessflat_raw.insert(g_lessflat_raw.end(), &a[4 * i], &a[4 * i] + 4);
    }
    const int m = n ? orc_voxelgrid(a.data(), n, 16, 3, leaf, b.data()) : 0;
    out.points.resize(m);
    for (int i = 0; i < m; ++i) out.points[i].x = b[4 * i], out.points[i].y = b[4 * i + 1], out.points[i].z = b[4 * i + 2], out.points[i].intensity = b[4 * i + 3];
    out.width = (uint32_t)m, out.height = 1;
  }
};
}  // namespace pcl

typedef pcl::PointXYZI PointType;  // parameters.h_ouster
using std::atan2;
using std::cos;
using std::sin;
#define ROS_BREAK() abort()

#include "globals.inc"
#include "remove.inc"

static const float* g_in = nullptr;
static int g_n = 0, g_stride_f = 4;
static void ref_fill(pcl::PointCloud<PointType>& c) {  // stands in for pcl::fromROSMsg(*laserCloudMsg, laserCloudIn)
  c.points.resize(g_n);
  for (int i = 0; i < g_n; ++i) {
    const float* p = g_in + (size_t)i * g_stride_f;
    c.points[i].x = p[0], c.points[i].y = p[1], c.points[i].z = p[2], c.points[i].intensity = g_stride_f > 3 ? p[3] : 0.f;
  }
  c.width = (uint32_t)g_n, c.height = 1;
}

struct RefScanRegOut {
  pcl::PointCloud<PointType> cloud, sharp, less_sharp, flat, less_flat;
  std::vector<int> start, end;
};
static RefScanRegOut g_out;

static void run_body() {
#include "body.inc"
  g_out.cloud = *laserCloud;
  g_out.sharp = cornerPointsSharp, g_out.less_sharp = cornerPointsLessSharp;
  g_out.flat = surfPointsFlat, g_out.less_flat = surfPointsLessFlat;
  g_out.start = scanStartInd, g_out.end = scanEndInd;
}

static int copy_out(const pcl::PointCloud<PointType>& c, float* out, int cap) {
  const int n = (int)c.points.size();
  for (int i = 0; i < n && i < cap; ++i) out[4 * i] = c.points[i].x, out[4 * i + 1] = c.points[i].y, out[4 * i + 2] = c.points[i].z, out[4 * i + 3] = c.points[i].intensity;
  return n;
}

// Runs the reference front end on n points (xyz[i] at stride_bytes).  Every output buffer holds cap points (4 floats each);
// counts[0..5] = sizes of cloud, sharp, less_sharp, flat, less_flat (after the VoxelGrid stand-in), less_flat_raw (before);
// curvature / label are per ring-ordered point; ring_start / ring_end = scanStartInd / scanEndInd (64 entries).
extern "C" int ref_scanreg(const float* xyz, int n, int stride_bytes, int n_scans, double minimum_range, int cap, float* cloud, float* curvature,
                           int32_t* label, float* sharp, float* less_sharp, float* flat, float* less_flat, float* less_flat_raw,
                           int32_t* ring_start, int32_t* ring_end, int32_t* counts) {
  g_in = xyz, g_n = n, g_stride_f = stride_bytes / 4;
  N_SCANS = n_scans, MINIMUM_RANGE = minimum_range;
  pcl::g_lessflat_raw.clear();
  g_out = RefScanRegOut();
  if (n > 400000) return -1;  // the reference's work arrays (:104-113)
  run_body();
  counts[0] = copy_out(g_out.cloud, cloud, cap);
  counts[1] = copy_out(g_out.sharp, sharp, cap);
  counts[2] = copy_out(g_out.less_sharp, less_sharp, cap);
  counts[3] = copy_out(g_out.flat, flat, cap);
  counts[4] = copy_out(g_out.less_flat, less_flat, cap);
  counts[5] = (int)(pcl::g_lessflat_raw.size() / 4);
  for (int i = 0; i < counts[5] && i < cap; ++i) memcpy(less_flat_raw + 4 * i, &pcl::g_lessflat_raw[4 * i], 16);
  for (int i = 0; i < counts[0] && i < cap; ++i) curvature[i] = cloudCurvature[i], label[i] = cloudLabel[i];
  for (int i = 0; i < n_scans && i < 64; ++i) ring_start[i] = g_out.start[i], ring_end[i] = g_out.end[i];
  return 0;
}
