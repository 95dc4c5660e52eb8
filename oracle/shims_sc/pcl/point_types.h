// Stand-in for <pcl/point_types.h>: the point type ScanContext reads (SCPointType = pcl::PointXYZI, x / y / z only).
// TEST INFRASTRUCTURE (oracle/_ref/libref_scancontext.so).
#pragma once
namespace pcl {
struct alignas(16) PointXYZI {
  float x = 0.f, y = 0.f, z = 0.f, data3 = 1.f;
  float intensity = 0.f, pad[3] = {0.f, 0.f, 0.f};
};
}  // namespace pcl
