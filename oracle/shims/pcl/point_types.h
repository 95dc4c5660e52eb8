// Stand-in for <pcl/point_types.h> (PCL is not installed here): only the point types ikd_Tree.cpp instantiates
// (ikd_Tree.cpp:2347-2356), with PCL's memory layout (parameters.h_ouster:121-124: PointXYZ 16 B, PointXYZI 32 B).
// TEST INFRASTRUCTURE for oracle/_ref/libref_ikd.so.
#pragma once
#include <string.h>

#include <vector>

#include <Eigen/StdVector>

namespace pcl {
struct alignas(16) PointXYZ {
  float x, y, z, data3;
  PointXYZ(float px = 0.f, float py = 0.f, float pz = 0.f) : x(px), y(py), z(pz), data3(1.f) {}
};
struct alignas(16) PointXYZI {
  float x, y, z, data3;
  float intensity, pad[3];
  PointXYZI(float px = 0.f, float py = 0.f, float pz = 0.f) : x(px), y(py), z(pz), data3(1.f), intensity(0.f), pad{0, 0, 0} {}
};
struct alignas(16) PointXYZINormal {
  float x, y, z, data3;
  float normal_x, normal_y, normal_z, data_n3;
  float intensity, curvature, pad[2];
  PointXYZINormal(float px = 0.f, float py = 0.f, float pz = 0.f)
      : x(px), y(py), z(pz), data3(1.f), normal_x(0), normal_y(0), normal_z(0), data_n3(0), intensity(0), curvature(0), pad{0, 0} {}
};
}  // namespace pcl
