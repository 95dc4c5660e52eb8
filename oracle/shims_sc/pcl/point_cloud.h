// Stand-in for <pcl/point_cloud.h>: a point vector with the members Scancontext.cpp touches.  TEST INFRASTRUCTURE.
#pragma once
#include <vector>
namespace pcl {
template <typename PointT>
struct PointCloud {
  std::vector<PointT> points;
  size_t size() const { return points.size(); }
};
}  // namespace pcl
