// ilsm_oracle_frontend.cpp -- CPU ORACLE (TEST INFRASTRUCTURE) for the front end: range/intensity image
// projection (image_handler.h_ouster:103-140), LOAM feature extraction (scanRegistration.cpp:152-186, 244-412,
// 427-589, upstream A-LOAM publish-once semantics) and PCL VoxelGrid (un-vendored, PCL 1.10 restated).
//
// PARITY STATUS: the feature extraction is PINNED against the reference's own code -- scanRegistration.cpp:227-589 cut out
// of its ROS node (oracle/patches/scanreg_extract.py) and compiled into oracle/_ref/libref_scanreg.so: ring-ordered cloud,
// relTime, ring bounds, curvature, labels and the four feature clouds agree bit for bit on natural frames
// (tests/test_ref_scanreg_cpu.py, tests/golden/scanreg_reference.npz).  Unpinned: PCL's VoxelGrid (un-vendored; the compiled
// reference fragment runs on THIS file's orc_voxelgrid).  The image projection is pinned the same way
// (image_handler.h_ouster:113-139 -> libref_imagehandler.so, tests/test_ref_imagehandler_cpu.py).  Decisions where the reference leaves the
// result unspecified are fixed and documented here:
//   * std::sort on curvature is unstable: ties are ordered by (curvature, point index) ascending (differs from libstdc++'s
//     introsort only under massive ties -- a perfectly regular synthetic cylinder; tested).
//   * PCL VoxelGrid sorts voxel records with an unstable sort: points of one voxel are accumulated in ascending
//     point index (float accumulators, like pcl::CentroidPoint).
//   * atan()/sqrt() in the ring formula resolve to the float overloads when <math.h> is included through
//     tf/LinearMath (ROS), to the double ones otherwise; the oracle evaluates atan in double and rounds to float
//     (= a correctly rounded atanf), which both libm variants agree with except in last-ulp cases.
#include <algorithm>
#include <cmath>
#include <cstdint>
#include <cstring>
#include <vector>

#define ORC_API extern "C" __attribute__((visibility("default")))

namespace {
inline const float* pt(const float* base, int stride_f, int i) { return base + (size_t)i * stride_f; }
}  // namespace

// ------------------------------------------------------------------------------------------------
// ImageHandler::cloud_handler (image_handler.h_ouster:103-140)
// intensity sits at float offset `ioff` (4 for pcl::PointXYZI 32-byte points, 3 for packed xyzi)
// ------------------------------------------------------------------------------------------------
ORC_API void orc_project(const float* cloud, int H, int W, int stride_bytes, int ioff, uint8_t* image_range,
                         uint8_t* image_intensity, float* cloud_track_xyzi) {
  const int sf = stride_bytes / 4;
  for (int u = 0; u < H; ++u)
    for (int v = 0; v < W; ++v) {
      const float* p = pt(cloud, sf, u * W + v);
      float range = std::sqrt(p[0] * p[0] + p[1] * p[1] + p[2] * p[2]);
      float intensity = p[ioff];
      intensity = std::min(intensity, 255.0f);
      float r20 = std::min(range * 20, 255.0f);
      image_range[u * W + v] = (uint8_t)(int32_t)r20;            // float -> uint8 conversion (truncation)
      image_intensity[u * W + v] = (uint8_t)(int32_t)intensity;  // x86 cvttss2si then low byte
      float* o = cloud_track_xyzi + 4 * (size_t)(u * W + v);
      if (range >= 0.1) {  // float compared against the double literal 0.1
        o[0] = p[0], o[1] = p[1], o[2] = p[2], o[3] = intensity;
      } else {
        o[0] = o[1] = o[2] = o[3] = 0.f;
      }
    }
}

// ------------------------------------------------------------------------------------------------
// pcl::VoxelGrid<PointXYZI>::applyFilter, PCL 1.10 (downsample_all_data = true, min_points_per_voxel = 0)
// in: n points xyzi (float4-like, stride), out: <= n centroids xyzi (packed 4 floats), returns count
// ------------------------------------------------------------------------------------------------
ORC_API int orc_voxelgrid(const float* in, int n, int stride_bytes, int ioff, float leaf, float* out_xyzi) {
  const int sf = stride_bytes / 4;
  if (n <= 0) return 0;
  float mn[3] = {INFINITY, INFINITY, INFINITY}, mx[3] = {-INFINITY, -INFINITY, -INFINITY};
  for (int i = 0; i < n; ++i) {
    const float* p = pt(in, sf, i);
    if (!std::isfinite(p[0]) || !std::isfinite(p[1]) || !std::isfinite(p[2])) continue;
    for (int a = 0; a < 3; ++a) mn[a] = std::min(mn[a], p[a]), mx[a] = std::max(mx[a], p[a]);
  }
  const float inv = 1.0f / leaf;
  int min_b[3], max_b[3], div_b[3];
  for (int a = 0; a < 3; ++a) {
    min_b[a] = (int)std::floor(mn[a] * inv);
    max_b[a] = (int)std::floor(mx[a] * inv);
    div_b[a] = max_b[a] - min_b[a] + 1;
  }
  const int mul[3] = {1, div_b[0], div_b[0] * div_b[1]};
  struct Rec {
    int64_t idx;
    int32_t pi;
  };
  std::vector<Rec> recs;
  recs.reserve(n);
  for (int i = 0; i < n; ++i) {
    const float* p = pt(in, sf, i);
    if (!std::isfinite(p[0]) || !std::isfinite(p[1]) || !std::isfinite(p[2])) continue;
    int ijk[3];
    for (int a = 0; a < 3; ++a) ijk[a] = (int)(std::floor(p[a] * inv) - (float)min_b[a]);
    recs.push_back({(int64_t)ijk[0] * mul[0] + (int64_t)ijk[1] * mul[1] + (int64_t)ijk[2] * mul[2], i});
  }
  std::stable_sort(recs.begin(), recs.end(), [](const Rec& a, const Rec& b) { return a.idx < b.idx; });
  int nout = 0;
  size_t k = 0;
  while (k < recs.size()) {
    size_t e = k;
    float sx = 0, sy = 0, sz = 0, si = 0;
    while (e < recs.size() && recs[e].idx == recs[k].idx) {
      const float* p = pt(in, sf, recs[e].pi);
      sx += p[0], sy += p[1], sz += p[2], si += p[ioff];
      ++e;
    }
    const float cnt = (float)(e - k);
    float* o = out_xyzi + 4 * (size_t)nout++;
    o[0] = sx / cnt, o[1] = sy / cnt, o[2] = sz / cnt, o[3] = si / cnt;
    k = e;
  }
  return nout;
}

// ------------------------------------------------------------------------------------------------
// scanRegistration.cpp laserCloudHandler (N_SCANS = 64 path)
// ------------------------------------------------------------------------------------------------
struct OrcFeatureCounts {
  int32_t n_cloud;  // ring-ordered cloud size (after min-range + ring filter)
  int32_t n_sharp, n_less_sharp, n_flat, n_less_flat;
  int32_t ring_start[64], ring_end[64];  // scanStartInd / scanEndInd
};

// outputs (all sized for n input points):
//   cloud_xyzi   : ring-ordered cloud, packed 4 floats, intensity = scanID + 0.1*relTime
//   curvature, label, picked(sort order: sort_ind) per ring-ordered point
//   src_index    : index of each ring-ordered point in the input cloud
//   sharp_idx / less_sharp_idx / flat_idx : indices into the ring-ordered cloud, in reference push order
//   less_flat_xyzi : the per-ring VoxelGrid(0.2)-filtered less-flat cloud, rings concatenated
ORC_API void orc_extract_features(const float* in, int n, int stride_bytes, float min_range, float* cloud_xyzi,
                                  float* curvature, int32_t* label, int32_t* src_index, int32_t* sharp_idx,
                                  int32_t* less_sharp_idx, int32_t* flat_idx, float* less_flat_xyzi,
                                  OrcFeatureCounts* counts) {
  const int sf = stride_bytes / 4;
  const int N_SCANS = 64;
  const double scanPeriod = 0.1;
  std::memset(counts, 0, sizeof(*counts));
  // removeClosedPointCloud (scanRegistration.cpp:152-186)
  std::vector<int> kept;
  kept.reserve(n);
  for (int i = 0; i < n; ++i) {
    const float* p = pt(in, sf, i);
    if (p[0] * p[0] + p[1] * p[1] + p[2] * p[2] < min_range * min_range) continue;
    kept.push_back(i);
  }
  int cloudSize = (int)kept.size();
  if (cloudSize == 0) return;
  const float* p0 = pt(in, sf, kept[0]);
  const float* pN = pt(in, sf, kept[cloudSize - 1]);
  float startOri = -std::atan2(p0[1], p0[0]);
  float endOri = (float)(-std::atan2(pN[1], pN[0]) + 2 * M_PI);
  if (endOri - startOri > 3 * M_PI)
    endOri = (float)(endOri - 2 * M_PI);
  else if (endOri - startOri < M_PI)
    endOri = (float)(endOri + 2 * M_PI);
  bool halfPassed = false;
  std::vector<std::vector<int>> ring_src(N_SCANS);
  std::vector<std::vector<float>> ring_int(N_SCANS);
  for (int c = 0; c < cloudSize; ++c) {
    const float* p = pt(in, sf, kept[c]);
    const float x = p[0], y = p[1], z = p[2];
    // correctly rounded float atan of (z / sqrtf(x^2+y^2)), * 180 in float, / M_PI in double, stored float
    const float ratio = z / std::sqrt(x * x + y * y);
    const float at = (float)std::atan((double)ratio);
    const float angle = (float)((at * 180) / M_PI);
    int scanID = int((angle + 22.5) * 1.41 + 0.5) - 1;
    if (scanID > (N_SCANS - 1) || scanID < 0) continue;
    float ori = -std::atan2(y, x);
    if (!halfPassed) {
      if (ori < startOri - M_PI / 2)
        ori = (float)(ori + 2 * M_PI);
      else if (ori > startOri + M_PI * 3 / 2)
        ori = (float)(ori - 2 * M_PI);
      if (ori - startOri > M_PI) halfPassed = true;
    } else {
      ori = (float)(ori + 2 * M_PI);
      if (ori < endOri - M_PI * 3 / 2)
        ori = (float)(ori + 2 * M_PI);
      else if (ori > endOri + M_PI / 2)
        ori = (float)(ori - 2 * M_PI);
    }
    const float relTime = (ori - startOri) / (endOri - startOri);
    ring_src[scanID].push_back(kept[c]);
    ring_int[scanID].push_back((float)(scanID + scanPeriod * relTime));
  }
  // concatenate rings
  int N = 0;
  for (int r = 0; r < N_SCANS; ++r) {
    counts->ring_start[r] = N + 5;
    for (size_t j = 0; j < ring_src[r].size(); ++j) {
      const float* p = pt(in, sf, ring_src[r][j]);
      float* o = cloud_xyzi + 4 * (size_t)N;
      o[0] = p[0], o[1] = p[1], o[2] = p[2], o[3] = ring_int[r][j];
      src_index[N] = ring_src[r][j];
      ++N;
    }
    counts->ring_end[r] = N - 6;
  }
  counts->n_cloud = N;
  for (int i = 0; i < N; ++i) curvature[i] = 0.f, label[i] = 0;
  std::vector<int> picked(N, 0), sortInd(N);
  auto P = [&](int i, int a) { return cloud_xyzi[4 * (size_t)i + a]; };
  for (int i = 5; i < N - 5; ++i) {
    float d[3];
    for (int a = 0; a < 3; ++a)
      d[a] = P(i - 5, a) + P(i - 4, a) + P(i - 3, a) + P(i - 2, a) + P(i - 1, a) - 10 * P(i, a) + P(i + 1, a) +
             P(i + 2, a) + P(i + 3, a) + P(i + 4, a) + P(i + 5, a);
    curvature[i] = d[0] * d[0] + d[1] * d[1] + d[2] * d[2];
    sortInd[i] = i;
  }
  auto gap2 = [&](int a, int b) {
    float dx = P(a, 0) - P(b, 0), dy = P(a, 1) - P(b, 1), dz = P(a, 2) - P(b, 2);
    return dx * dx + dy * dy + dz * dz;
  };
  auto mark = [&](int ind) {
    picked[ind] = 1;
    for (int l = 1; l <= 5; l++) {
      if (gap2(ind + l, ind + l - 1) > 0.05) break;
      picked[ind + l] = 1;
    }
    for (int l = -1; l >= -5; l--) {
      if (gap2(ind + l, ind + l + 1) > 0.05) break;
      picked[ind + l] = 1;
    }
  };
  int ns = 0, nls = 0, nf = 0, nlf = 0;
  std::vector<float> ring_lf, ring_ds;
  for (int r = 0; r < N_SCANS; ++r) {
    const int S = counts->ring_start[r], E = counts->ring_end[r];
    if (E - S < 6) continue;
    ring_lf.clear();
    for (int j = 0; j < 6; ++j) {
      const int sp = S + (E - S) * j / 6, ep = S + (E - S) * (j + 1) / 6 - 1;
      std::sort(sortInd.begin() + sp, sortInd.begin() + ep + 1, [&](int a, int b) {
        return curvature[a] < curvature[b] || (curvature[a] == curvature[b] && a < b);
      });
      int largest = 0;
      for (int k = ep; k >= sp; --k) {
        const int ind = sortInd[k];
        if (picked[ind] == 0 && curvature[ind] > 0.1) {
          ++largest;
          if (largest <= 2) {
            label[ind] = 2;
            sharp_idx[ns++] = ind;
            less_sharp_idx[nls++] = ind;
          } else if (largest <= 20) {
            label[ind] = 1;
            less_sharp_idx[nls++] = ind;
          } else {
            break;
          }
          mark(ind);
        }
      }
      int smallest = 0;
      for (int k = sp; k <= ep; ++k) {
        const int ind = sortInd[k];
        if (picked[ind] == 0 && curvature[ind] < 0.1) {
          label[ind] = -1;
          flat_idx[nf++] = ind;
          ++smallest;
          if (smallest >= 4) break;  // before marking (scanRegistration.cpp:530-533)
          mark(ind);
        }
      }
      for (int k = sp; k <= ep; ++k)
        if (label[k] <= 0) {
          for (int a = 0; a < 4; ++a) ring_lf.push_back(P(k, a));
        }
    }
    ring_ds.resize(ring_lf.size());
    const int m = orc_voxelgrid(ring_lf.data(), (int)ring_lf.size() / 4, 16, 3, 0.2f, ring_ds.data());
    std::memcpy(less_flat_xyzi + 4 * (size_t)nlf, ring_ds.data(), sizeof(float) * 4 * (size_t)m);
    nlf += m;
  }
  counts->n_sharp = ns, counts->n_less_sharp = nls, counts->n_flat = nf, counts->n_less_flat = nlf;
}
