// ref_imagehandler.cpp -- TEST INFRASTRUCTURE.  Runs the reference's own image projection loop
// (src/image_handler.h_ouster:113-139, cut out of ImageHandler::cloud_handler by oracle/patches/imagehandler_extract.py) on a
// caller-supplied organised cloud: the 8-bit range and intensity images and the cloud_track copy.  Stand-ins (not the
// reference): cv::Mat as a byte matrix with at<uint8_t>(u, v), pcl::PointCloud as a point vector.
// Built only into oracle/_ref/libref_imagehandler.so (git-ignored); nothing in the product path links it.
#include <algorithm>
#include <cmath>
#include <cstdint>
#include <memory>
#include <vector>

namespace cv {
struct Mat {
  int rows = 0, cols = 0;
  std::vector<uint8_t> d;
  Mat() {}
  Mat(int r, int c) : rows(r), cols(c), d((size_t)r * c, 0) {}
  template <typename T>
  T& at(int u, int v) { return reinterpret_cast<T&>(d[(size_t)u * cols + v]); }
};
}  // namespace cv
struct PointOuster {  // the fields the loop reads (parameters.h_ouster: x, y, z, intensity, ...)
  float x, y, z, intensity;
};
struct PointXYZI {
  float x, y, z, intensity;
};
typedef PointXYZI PointType;
template <typename P>
struct Cloud {
  std::vector<P> points;
};

// image_range / image_intensity row-major [H][W]; track packed xyzi [H * W]
extern "C" void ref_project(const float* xyzi, int H, int W, int stride_bytes, uint8_t* image_range_out, uint8_t* image_intensity_out,
                            float* track_out) {
  const int IMAGE_HEIGHT = H, IMAGE_WIDTH = W, sf = stride_bytes / 4;
  std::shared_ptr<Cloud<PointOuster>> laser_cloud(new Cloud<PointOuster>());
  std::shared_ptr<Cloud<PointType>> cloud_track(new Cloud<PointType>());
  laser_cloud->points.resize((size_t)H * W), cloud_track->points.resize((size_t)H * W);  // :35-36 sizes cloud_track once
  for (size_t i = 0; i < (size_t)H * W; ++i) {
    const float* p = xyzi + i * sf;
    laser_cloud->points[i] = PointOuster{p[0], p[1], p[2], p[3]};
    cloud_track->points[i] = PointType{-1.f, -1.f, -1.f, -1.f};  // every entry is overwritten by the loop
  }
  cv::Mat image_range(H, W), image_intensity(H, W), image_ambient(H, W);
#include "loop.inc"
  for (size_t i = 0; i < (size_t)H * W; ++i) {
    image_range_out[i] = image_range.d[i], image_intensity_out[i] = image_intensity.d[i];
    const PointType& t = cloud_track->points[i];
    track_out[4 * i] = t.x, track_out[4 * i + 1] = t.y, track_out[4 * i + 2] = t.z, track_out[4 * i + 3] = t.intensity;
  }
}
