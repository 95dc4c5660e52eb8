// ref_scancontext.cpp -- TEST INFRASTRUCTURE.  Drives the reference's own ScanContext code -- src/Scancontext.cpp and
// include/Scancontext.h, compiled UNMODIFIED from where they lie (this file needs no cuts or repairs) together with the
// vendored nanoflann and KDTreeVectorOfVectorsAdaptor -- so that the oracle's restatement of the descriptor, the two keys,
// the column-shift distance and detectLoopClosureID can be compared with the reference code.
// Stand-ins (not the reference): Eigen::MatrixXd (oracle/shims_sc/Eigen/Dense: element access and views are exact, the
// three reductions norm / dot / mean run left to right where Eigen vectorises), pcl::PointCloud, empty OpenCV / cv_bridge
// headers.  Built only into oracle/_ref/libref_scancontext.so (git-ignored); nothing in the product path links it.
#include <fcntl.h>
#include <unistd.h>

#include <cstdint>
#include <cstdio>
#include <cstring>

#include "Scancontext.h"

static Eigen::MatrixXd to_mat(const double* row_major) {  // [20][60] row-major -> the reference's (ring, sector) matrix
  Eigen::MatrixXd m(20, 60);
  for (int r = 0; r < 20; ++r)
    for (int c = 0; c < 60; ++c) m(r, c) = row_major[r * 60 + c];
  return m;
}

// SCManager::makeScancontext on n points (xyz at stride_bytes): desc_out row-major [20][60]
extern "C" __attribute__((visibility("default"))) void ref_sc_make(const float* pts, int n, int stride_bytes, double* desc_out) {
  pcl::PointCloud<SCPointType> cloud;
  cloud.points.resize(n);
  const int sf = stride_bytes / 4;
  for (int i = 0; i < n; ++i) cloud.points[i].x = pts[(size_t)i * sf], cloud.points[i].y = pts[(size_t)i * sf + 1], cloud.points[i].z = pts[(size_t)i * sf + 2];
  SCManager m;
  const Eigen::MatrixXd d = m.makeScancontext(cloud);
  for (int r = 0; r < 20; ++r)
    for (int c = 0; c < 60; ++c) desc_out[r * 60 + c] = d(r, c);
}

// makeRingkeyFromScancontext / makeSectorkeyFromScancontext
extern "C" __attribute__((visibility("default"))) void ref_sc_keys(const double* desc, double* ring_key20, double* sector_key60) {
  SCManager m;
  Eigen::MatrixXd d = to_mat(desc);
  const Eigen::MatrixXd rk = m.makeRingkeyFromScancontext(d), sk = m.makeSectorkeyFromScancontext(d);
  for (int r = 0; r < 20; ++r) ring_key20[r] = rk(r, 0);
  for (int c = 0; c < 60; ++c) sector_key60[c] = sk(0, c);
}

// distanceBtnScanContext(query, candidate) -> (distance, aligning shift)
extern "C" __attribute__((visibility("default"))) void ref_sc_distance(const double* q, const double* c, double* dist, int32_t* shift) {
  SCManager m;
  Eigen::MatrixXd a = to_mat(q), b = to_mat(c);
  const std::pair<double, int> r = m.distanceBtnScanContext(a, b);
  *dist = r.first, *shift = r.second;
}

// detectLoopClosureID with the database descs[0 .. n) pushed in order (the last one is the query, as in the node: the
// current keyframe is saved first, Scancontext.cpp:237-251).  Returns the loop id (-1: none) and the yaw difference.
extern "C" __attribute__((visibility("default"))) int ref_sc_detect(const double* descs, int n, float* yaw_diff_rad) {
  SCManager m;
  for (int i = 0; i < n; ++i) {
    Eigen::MatrixXd sc = to_mat(descs + (size_t)i * 1200);
    Eigen::MatrixXd ringkey = m.makeRingkeyFromScancontext(sc), sectorkey = m.makeSectorkeyFromScancontext(sc);
    m.polarcontexts_.push_back(sc);
    m.polarcontext_invkeys_.push_back(ringkey);
    m.polarcontext_vkeys_.push_back(sectorkey);
    m.polarcontext_invkeys_mat_.push_back(eig2stdvec(ringkey));
  }
  // detectLoopClosureID reports on std::cout (every line ends in std::endl, i.e. is flushed).  File descriptor 1 is pointed at
  // /dev/null for the call; std::cout itself is left alone -- with libstdc++ linked statically into several libraries of one
  // process its objects are unified by the dynamic linker and must not be reconfigured from here.
  fflush(stdout);
  const int saved = dup(1), nul = open("/dev/null", O_WRONLY);
  if (saved >= 0 && nul >= 0) dup2(nul, 1);
  const std::pair<int, float> r = m.detectLoopClosureID();
  fflush(stdout);
  if (saved >= 0) dup2(saved, 1), close(saved);
  if (nul >= 0) close(nul);
  *yaw_diff_rad = r.second;
  return r.first;
}
