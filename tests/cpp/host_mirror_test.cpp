// host_mirror_test.cpp -- parity of the C++ host side (include/ilsm.hpp: KdTreeFLANN, VoxelGrid, KD_TREE, SCManager,
// ImageHandler, ScanRegistration, ScanToMapRegistration) against the CPU oracle, written the way a unit test of the
// reference's own objects would read.  Links libilsm_cuda.so (the product) and libilsm_oracle.so (the checker: test
// infrastructure only).  Built and run by tests/test_cpp_host.py on the GPU box; exit code = number of failed checks.
#include <algorithm>
#include <cstdio>
#include <cstdlib>
#include <random>

#include "ilsm.hpp"

// ---- oracle (oracle/ilsm_oracle*.cpp) ----
struct OrcSolveSummary {
  int32_t termination, iterations, num_successful, num_unsuccessful;
  double initial_cost, final_cost;
  int32_t num_evals, pad;
};
struct OrcFeatureCounts {
  int32_t n_cloud, n_sharp, n_less_sharp, n_flat, n_less_flat;
  int32_t ring_start[64], ring_end[64];
};
extern "C" {
void orc_knn_brute(const float* map, int n, int map_stride_bytes, const float* q, int nq, int q_stride_bytes, int k, int32_t* idx,
                   float* d2);
int orc_voxelgrid(const float* in, int n, int stride_bytes, int ioff, float leaf, float* out_xyzi);
int orc_ikd_add_points(const float* existing, int n_old, const float* add, int n_add, float ds, int downsample, float* out_xyz,
                       int out_cap);
void orc_project(const float* cloud, int H, int W, int stride_bytes, int ioff, uint8_t* image_range, uint8_t* image_intensity,
                 float* cloud_track_xyzi);
void orc_extract_features(const float* in, int n, int stride_bytes, float min_range, float* cloud_xyzi, float* curvature,
                          int32_t* label, int32_t* src_index, int32_t* sharp_idx, int32_t* less_sharp_idx, int32_t* flat_idx,
                          float* less_flat_xyzi, OrcFeatureCounts* counts);
int orc_register_aloam(const float* map_corner, int n_mc, const float* map_surf, int n_ms, int map_stride_bytes, const float* corner,
                       int nc, const float* surf, int ns, int stride_bytes, double qt[7], int outer, int max_iter,
                       OrcSolveSummary* summaries, int32_t* nfactors);
void orc_sc_make(const float* pts, int n, int stride_bytes, double* desc);
void orc_sc_topk(const double* db, int n, const double* q, int k, double* dist, int32_t* id, int32_t* shift);
}

using ilsm::PointXYZI;
typedef ilsm::PointCloud<PointXYZI> Cloud;

static int g_failed = 0;
#define EXPECT(cond)                                                        \
  do {                                                                      \
    if (!(cond)) {                                                          \
      std::printf("  FAILED %s:%d: %s\n", __FILE__, __LINE__, #cond);       \
      ++g_failed;                                                           \
    }                                                                       \
  } while (0)

static PointXYZI pt(float x, float y, float z, float i = 0.f) {
  PointXYZI p;
  p.x = x, p.y = y, p.z = z, p.intensity = i;
  return p;
}

// Box room [-12,12] x [-9,9] x [-1.6,3.4] with two interior pillars: analytic ray cast of an organised H x W frame.
static float cast(const double o[3], const double d[3]) {
  double best = 1e9;
  const double lo[3] = {-12, -9, -1.6}, hi[3] = {12, 9, 3.4};
  for (int a = 0; a < 3; ++a) {
    if (std::fabs(d[a]) < 1e-12) continue;
    const double t = ((d[a] > 0 ? hi[a] : lo[a]) - o[a]) / d[a];
    if (t > 0 && t < best) best = t;
  }
  const double pil[2][4] = {{4, 5, 2, 3}, {-6, -5, -4, -3}};  // x0,x1,y0,y1
  for (int k = 0; k < 2; ++k) {
    double t0 = 0, t1 = 1e9;
    bool ok = true;
    for (int a = 0; a < 2 && ok; ++a) {
      const double l = pil[k][2 * a], h = pil[k][2 * a + 1];
      if (std::fabs(d[a]) < 1e-12) {
        ok = o[a] >= l && o[a] <= h;
        continue;
      }
      double ta = (l - o[a]) / d[a], tb = (h - o[a]) / d[a];
      if (ta > tb) std::swap(ta, tb);
      t0 = std::max(t0, ta), t1 = std::min(t1, tb);
      ok = t0 <= t1;
    }
    if (ok && t0 > 0 && t0 < best) best = t0;
  }
  return (float)best;
}

static Cloud make_frame(int H, int W, const double origin[3], double yaw, unsigned seed, double half_fov_deg = 22.0) {
  std::mt19937 rng(seed);
  std::normal_distribution<float> noise(0.f, 0.005f);
  std::uniform_real_distribution<float> inten(0.f, 255.f);
  Cloud c;
  c.points.resize((size_t)H * W);
  for (int u = 0; u < H; ++u) {
    const double el = (half_fov_deg - 2.0 * half_fov_deg * u / (H - 1)) * M_PI / 180.0;
    for (int v = 0; v < W; ++v) {
      // columns offset by a fraction of a step: no azimuth sits exactly on the +-pi/2 wrap thresholds of
      // scanRegistration.cpp:330-352, where the last ulp of atan2f (libm vs CUDA) would decide relTime
      const double az = yaw + 2.0 * M_PI * (v + 0.37) / W;
      const double dl[3] = {std::cos(el) * std::cos(az), std::cos(el) * std::sin(az), std::sin(el)};
      const float r = cast(origin, dl) + noise(rng);
      // points in the SENSOR frame (sensor axes = world axes rotated by yaw about z, sensor at `origin`)
      const double ds[3] = {std::cos(el) * std::cos(az - yaw), std::cos(el) * std::sin(az - yaw), std::sin(el)};
      c.points[(size_t)u * W + v] = (v % 97 == 0) ? pt(0, 0, 0, 0) : pt((float)(r * ds[0]), (float)(r * ds[1]), (float)(r * ds[2]), inten(rng));
    }
  }
  c.width = W, c.height = H;
  return c;
}

static void test_kdtree_flann() {
  std::printf("KdTreeFLANN::setInputCloud / nearestKSearch vs brute force\n");
  std::mt19937 rng(1);
  std::uniform_real_distribution<float> u(-20.f, 20.f);
  Cloud::Ptr map(new Cloud());
  for (int i = 0; i < 20000; ++i) map->push_back(pt(u(rng), u(rng), 0.2f * u(rng)));
  map->points[77] = map->points[76];  // exact tie: lower index first
  Cloud q;
  for (int i = 0; i < 600; ++i) q.push_back(pt(u(rng), u(rng), 0.2f * u(rng)));
  q.points[0] = map->points[76];
  ilsm::KdTreeFLANN<PointXYZI> kdtree;
  kdtree.setInputCloud(map);
  EXPECT(kdtree.size() == 20000);
  std::vector<int32_t> ri(600 * 5);
  std::vector<float> rd(600 * 5);
  orc_knn_brute(&map->points[0].x, 20000, 32, &q.points[0].x, 600, 32, 5, ri.data(), rd.data());
  std::vector<int> idx;
  std::vector<float> d2;
  for (int i = 0; i < 60; ++i) {  // the reference's one-point-at-a-time form
    EXPECT(kdtree.nearestKSearch(q.points[i], 5, idx, d2) == 5);
    for (int k = 0; k < 5; ++k) EXPECT(idx[k] == ri[5 * i + k] && d2[k] == rd[5 * i + k]);
  }
  EXPECT(kdtree.nearestKSearch(q, 5, idx, d2) == 600);  // batched
  bool same = true;
  for (size_t j = 0; j < idx.size(); ++j) same = same && idx[j] == ri[j] && d2[j] == rd[j];
  EXPECT(same);
  Cloud tiny;  // fewer points than k: PCL returns what exists
  tiny.push_back(pt(0, 0, 0)), tiny.push_back(pt(1, 0, 0));
  kdtree.setInputCloud(tiny);
  EXPECT(kdtree.nearestKSearch(pt(0.9f, 0, 0), 5, idx, d2) == 2 && idx[0] == 1 && idx[1] == 0);
}

static void test_voxelgrid() {
  std::printf("VoxelGrid::filter vs PCL restatement\n");
  std::mt19937 rng(2);
  std::uniform_real_distribution<float> u(-8.f, 8.f);
  Cloud::Ptr in(new Cloud());
  for (int i = 0; i < 6000; ++i) in->push_back(pt(u(rng), u(rng), 0.1f * u(rng), (float)(i % 64) + 0.05f));
  ilsm::VoxelGrid<PointXYZI> f;
  f.setLeafSize(0.4f, 0.4f, 0.4f);
  f.setInputCloud(in);
  Cloud out;
  f.filter(out);
  std::vector<float> ref(6000 * 4);
  const int n = orc_voxelgrid(&in->points[0].x, 6000, 32, 4, 0.4f, ref.data());
  EXPECT((int)out.size() == n && n > 1000 && n < 6000);
  bool same = (int)out.size() == n;
  for (int i = 0; same && i < n; ++i)
    same = out[i].x == ref[4 * i] && out[i].y == ref[4 * i + 1] && out[i].z == ref[4 * i + 2] && out[i].intensity == ref[4 * i + 3];
  EXPECT(same);
}

static void test_ikd_tree() {
  std::printf("KD_TREE::Build / Nearest_Search / Add_Points(downsample) / flatten vs sequential restatement\n");
  std::mt19937 rng(3);
  std::uniform_real_distribution<float> u(-6.f, 6.f);
  ilsm::KD_TREE::PointVector base, add;
  for (int i = 0; i < 3000; ++i) base.push_back(ilsm::KD_TREE::PointType(u(rng), u(rng), 0.05f * u(rng)));
  for (int i = 0; i < 1500; ++i) add.push_back(ilsm::KD_TREE::PointType(1.2f * u(rng), 1.2f * u(rng), 0.05f * u(rng)));
  ilsm::KD_TREE ikdtree(0.5f, 0.6f, 0.4f);
  ikdtree.Build(base);
  EXPECT(ikdtree.size() == 3000);
  ilsm::KD_TREE::PointVector near;
  std::vector<float> dist;
  std::vector<int32_t> ri(5);
  std::vector<float> rd(5);
  for (int i = 0; i < 40; ++i) {
    ilsm::KD_TREE::PointType p(u(rng), u(rng), 0.05f * u(rng));
    ikdtree.Nearest_Search(p, 5, near, dist);
    orc_knn_brute(&base[0].x, 3000, 12, &p.x, 1, 12, 5, ri.data(), rd.data());
    EXPECT(near.size() == 5 && dist.size() == 5);
    for (int k = 0; k < 5 && k < (int)near.size(); ++k)
      EXPECT(dist[k] == rd[k] && near[k].x == base[ri[k]].x && near[k].y == base[ri[k]].y && near[k].z == base[ri[k]].z);
  }
  EXPECT(ikdtree.Add_Points(add, true) == 1500);
  std::vector<float> ref((3000 + 1500) * 3);
  const int n_ref = orc_ikd_add_points(&base[0].x, 3000, &add[0].x, 1500, 0.4f, 1, ref.data(), 4500);
  ilsm::KD_TREE::PointVector flat;
  ikdtree.flatten(ikdtree.Root_Node, flat, ilsm::NOT_RECORD);
  EXPECT((int)flat.size() == n_ref && ikdtree.size() == n_ref);
  auto key = [](const float* p) { return std::make_tuple(p[0], p[1], p[2]); };
  std::vector<std::tuple<float, float, float>> a, b;
  for (auto& p : flat) a.push_back(key(&p.x));
  for (int i = 0; i < n_ref; ++i) b.push_back(key(&ref[3 * i]));
  std::sort(a.begin(), a.end()), std::sort(b.begin(), b.end());
  EXPECT(a == b);
  // the search structure follows the insertion
  ilsm::KD_TREE::PointType p(0.1f, 0.2f, 0.f);
  ikdtree.Nearest_Search(p, 5, near, dist);
  orc_knn_brute(ref.data(), n_ref, 12, &p.x, 1, 12, 5, ri.data(), rd.data());
  for (int k = 0; k < 5 && k < (int)dist.size(); ++k) EXPECT(dist[k] == rd[k]);
}

static void test_image_handler_and_scan_registration() {
  std::printf("ImageHandler::cloud_handler and ScanRegistration::laserCloudHandler vs restatement\n");
  const int H = 64, W = 1024;
  const double origin[3] = {0.5, -0.3, 0.0};
  Cloud frame = make_frame(H, W, origin, 0.3, 11);
  ilsm::ImageHandler ih(H, W);
  ih.cloud_handler(frame);
  std::vector<uint8_t> rr(H * W), ri(H * W);
  std::vector<float> track(H * W * 4);
  orc_project(&frame.points[0].x, H, W, 32, 4, rr.data(), ri.data(), track.data());
  EXPECT(ih.image_range == rr && ih.image_intensity == ri);
  bool same = true;
  for (int i = 0; i < H * W; ++i)
    same = same && ih.cloud_track->points[i].x == track[4 * i] && ih.cloud_track->points[i].intensity == track[4 * i + 3];
  EXPECT(same);

  ilsm::ScanRegistration sr(0.3f);
  sr.laserCloudHandler(frame);
  const int n = H * W;
  std::vector<float> cloud(n * 4), curv(n), lflat(n * 4);
  std::vector<int32_t> label(n), src(n), sharp(n), lsharp(n), flat(n);
  OrcFeatureCounts cnt;
  orc_extract_features(&frame.points[0].x, n, 32, 0.3f, cloud.data(), curv.data(), label.data(), src.data(), sharp.data(),
                       lsharp.data(), flat.data(), lflat.data(), &cnt);
  EXPECT(sr.counts.n_cloud == cnt.n_cloud && sr.counts.n_sharp == cnt.n_sharp && sr.counts.n_less_sharp == cnt.n_less_sharp &&
         sr.counts.n_flat == cnt.n_flat && sr.counts.n_less_flat == cnt.n_less_flat);
  EXPECT(cnt.n_sharp > 20 && cnt.n_flat > 200 && cnt.n_less_flat > 1000);
  auto eq = [](const Cloud& c, const float* xyzi, const int32_t* idx, int m) {
    if ((int)c.size() != m) return false;
    for (int i = 0; i < m; ++i) {
      const float* s = xyzi + 4 * (size_t)(idx ? idx[i] : i);
      // xyz and the ring id bit-exact; relTime goes through atan2f whose last ulp differs between libm and CUDA
      if (c[i].x != s[0] || c[i].y != s[1] || c[i].z != s[2] || (int)c[i].intensity != (int)s[3] ||
          std::fabs(c[i].intensity - s[3]) > 1.6e-5f)
        return false;
    }
    return true;
  };
  EXPECT(eq(sr.laserCloud, cloud.data(), nullptr, cnt.n_cloud));
  EXPECT(eq(sr.cornerPointsSharp, cloud.data(), sharp.data(), cnt.n_sharp));
  EXPECT(eq(sr.cornerPointsLessSharp, cloud.data(), lsharp.data(), cnt.n_less_sharp));
  EXPECT(eq(sr.surfPointsFlat, cloud.data(), flat.data(), cnt.n_flat));
  EXPECT(eq(sr.surfPointsLessFlat, lflat.data(), nullptr, cnt.n_less_flat));
}

static void test_scan_to_map_registration() {
  std::printf("ScanToMapRegistration::align (2 x (association + ceres::Solve)) vs restatement\n");
  // map: the room's surfaces on a jittered 0.4 m lattice (surf) and its vertical edges / pillar edges (corner), world frame
  std::mt19937 rng(5);
  std::uniform_real_distribution<float> j(-0.03f, 0.03f);
  Cloud surfMap, cornerMap;
  for (float x = -12; x <= 12; x += 0.4f)
    for (float y = -9; y <= 9; y += 0.4f) surfMap.push_back(pt(x + j(rng), y + j(rng), -1.6f + 0.2f * j(rng)));
  for (float x = -12; x <= 12; x += 0.4f)
    for (float z = -1.6f; z <= 3.4f; z += 0.4f) {
      surfMap.push_back(pt(x + j(rng), -9 + 0.2f * j(rng), z + j(rng)));
      surfMap.push_back(pt(x + j(rng), 9 + 0.2f * j(rng), z + j(rng)));
    }
  for (float y = -9; y <= 9; y += 0.4f)
    for (float z = -1.6f; z <= 3.4f; z += 0.4f) {
      surfMap.push_back(pt(-12 + 0.2f * j(rng), y + j(rng), z + j(rng)));
      surfMap.push_back(pt(12 + 0.2f * j(rng), y + j(rng), z + j(rng)));
    }
  const float ex[8][2] = {{-12, -9}, {-12, 9}, {12, -9}, {12, 9}, {4, 2}, {5, 3}, {-6, -4}, {-5, -3}};
  for (int e = 0; e < 8; ++e)
    for (float z = -1.6f; z <= 3.4f; z += 0.1f) cornerMap.push_back(pt(ex[e][0] + 0.1f * j(rng), ex[e][1] + 0.1f * j(rng), z));
  // true pose and the stacks: map points seen from it, moved into the sensor frame
  const double yaw = 0.2, tt[3] = {0.6, -0.4, 0.1};
  const double qt_true[7] = {0, 0, std::sin(yaw / 2), std::cos(yaw / 2), tt[0], tt[1], tt[2]};
  auto to_sensor = [&](const PointXYZI& w) {
    const double dx = w.x - tt[0], dy = w.y - tt[1], dz = w.z - tt[2];
    return pt((float)(std::cos(yaw) * dx + std::sin(yaw) * dy) + 0.2f * j(rng), (float)(-std::sin(yaw) * dx + std::cos(yaw) * dy) + 0.2f * j(rng),
              (float)dz + 0.2f * j(rng));
  };
  Cloud cornerStack, surfStack;
  for (size_t i = 0; i < cornerMap.size(); i += 2) cornerStack.push_back(to_sensor(cornerMap[i]));
  for (size_t i = 0; i < surfMap.size(); i += 3) surfStack.push_back(to_sensor(surfMap[i]));
  double parameters[7] = {0, 0, std::sin(0.23 / 2), std::cos(0.23 / 2), 0.72, -0.31, 0.18};  // perturbed guess
  double ref[7];
  std::memcpy(ref, parameters, sizeof(ref));

  ilsm::ScanToMapRegistration<PointXYZI> reg;
  EXPECT(reg.align(cornerMap, surfMap, cornerStack, surfStack, parameters));
  OrcSolveSummary sum[2];
  int32_t nf[4];
  EXPECT(orc_register_aloam(&cornerMap.points[0].x, (int)cornerMap.size(), &surfMap.points[0].x, (int)surfMap.size(), 32,
                            &cornerStack.points[0].x, (int)cornerStack.size(), &surfStack.points[0].x, (int)surfStack.size(), 32, ref,
                            2, 4, sum, nf) == 2);
  EXPECT(reg.report.passes == 2);
  for (int p = 0; p < 2; ++p) {
    EXPECT(reg.report.pass[p].num_edge_factors == nf[2 * p] && reg.report.pass[p].num_plane_factors == nf[2 * p + 1]);
    EXPECT(reg.report.pass[p].termination == sum[p].termination && reg.report.pass[p].iterations == sum[p].iterations);
    EXPECT(std::fabs(reg.report.pass[p].final_cost - sum[p].final_cost) <= 1e-5 * sum[p].final_cost);  // north star: 1e-5 relative
  }
  EXPECT(nf[0] > 50 && nf[1] > 500);
  double dt = 0, dq = 0, err = 0;
  for (int i = 0; i < 3; ++i) dt = std::max(dt, std::fabs(parameters[4 + i] - ref[4 + i])), err = std::max(err, std::fabs(parameters[4 + i] - qt_true[4 + i]));
  for (int i = 0; i < 4; ++i) dq = std::max(dq, std::fabs(parameters[i] - ref[i]));
  EXPECT(dt < 1e-4 && 2 * dq < 1e-4);  // north star: 1e-4 m / 1e-4 rad
  EXPECT(err < 0.03);                  // and the frame is actually registered
  std::printf("  pose vs oracle: %.2e m, %.2e (quat); vs truth %.3f m\n", dt, dq, err);
  // the guard of laserMapping.cpp:624
  Cloud few;
  for (int i = 0; i < 5; ++i) few.push_back(cornerMap[i]);
  EXPECT(!reg.align(few, surfMap, cornerStack, surfStack, parameters));
}

static void test_loam_pipeline() {
  std::printf("LoamPipeline: synchronous vs pipelined mapping stage vs three stages (same poses, one / two calls later)\n");
  // an OS0-like +-45 deg sensor: the 64-ring formula keeps the beams within +-22.5 deg (scanRegistration.cpp:308-316), which
  // also keeps the less-flat cloud under the 16384 points one ilsm_slam_frame call accepts per feature cloud
  const int H = 64, W = 1024, F = 6;
  std::vector<Cloud> frames;
  for (int k = 0; k < F; ++k) {
    const double origin[3] = {0.5 + 0.15 * k, -0.3 + 0.02 * k, 0.0};
    frames.push_back(make_frame(H, W, origin, 0.3 + 0.01 * k, 100 + k, 45.0));
  }
  ilsm::LoamPipeline sync(0.4f, 0.8f, 0.3f, false), pipe(0.4f, 0.8f, 0.3f, true);  // default cube capacity (16384 points per cube)
  std::vector<std::vector<double>> mapped_sync, mapped_pipe;
  for (int k = 0; k < F; ++k) {
    sync.frame(frames[k]);
    pipe.frame(frames[k]);
    EXPECT(sync.has_mapped_pose && pipe.has_mapped_pose == (k > 0));
    for (int i = 0; i < 4; ++i) EXPECT(sync.q_odom[i] == pipe.q_odom[i]);
    for (int i = 0; i < 3; ++i) EXPECT(sync.t_odom[i] == pipe.t_odom[i]);
    mapped_sync.push_back({sync.q_map[0], sync.q_map[1], sync.q_map[2], sync.q_map[3], sync.t_map[0], sync.t_map[1], sync.t_map[2]});
    if (pipe.has_mapped_pose)
      mapped_pipe.push_back({pipe.q_map[0], pipe.q_map[1], pipe.q_map[2], pipe.q_map[3], pipe.t_map[0], pipe.t_map[1], pipe.t_map[2]});
  }
  EXPECT(pipe.flush());
  mapped_pipe.push_back({pipe.q_map[0], pipe.q_map[1], pipe.q_map[2], pipe.q_map[3], pipe.t_map[0], pipe.t_map[1], pipe.t_map[2]});
  EXPECT(!pipe.flush());
  EXPECT(mapped_pipe == mapped_sync);
  EXPECT(sync.stats.n_less_flat > 500);
  // ... and as three stages: odometry one call later, mapped pose two calls later, the same numbers
  ilsm::StagedLoamPipeline staged(0.4f, 0.8f, 0.3f);
  std::vector<std::vector<double>> mapped_staged(F);
  int n_mapped = 0;
  for (int k = 0; k < F + 2; ++k) {
    if (k < F) staged.push(frames[k]);
    else staged.drain();
    EXPECT(staged.odom_frame == (k >= 1 && k <= F ? k - 1 : -1) && staged.map_frame == (k >= 2 ? k - 2 : -1));
    if (staged.map_frame >= 0) {
      mapped_staged[staged.map_frame] = {staged.q_map[0], staged.q_map[1], staged.q_map[2], staged.q_map[3], staged.t_map[0], staged.t_map[1], staged.t_map[2]};
      ++n_mapped;
    }
  }
  EXPECT(n_mapped == F && mapped_staged == mapped_sync);
  // informational: the sensor moved by (0.75, 0.10) m in the world, i.e. (0.746, -0.126) m in the first sensor frame (yaw 0.3)
  std::printf("  mapped translation after %d frames: %.3f %.3f %.3f (sensor displacement in the map frame: 0.746 -0.126 0.000)\n", F,
              sync.t_map[0], sync.t_map[1], sync.t_map[2]);
}

static void test_scancontext() {
  std::printf("SCManager::makeAndSaveScancontextAndKeys / detectLoopClosureID vs restatement\n");
  ilsm::SCManager sc;
  std::mt19937 rng(9);
  std::uniform_real_distribution<float> ang(0.f, 6.2831853f), rad(1.f, 75.f), hgt(-1.5f, 4.f);
  std::vector<Cloud> scans;
  std::vector<double> db;
  auto rotate = [](const Cloud& c, double yaw) {
    Cloud o;
    for (auto& p : c.points) o.push_back(pt((float)(std::cos(yaw) * p.x - std::sin(yaw) * p.y), (float)(std::sin(yaw) * p.x + std::cos(yaw) * p.y), p.z));
    return o;
  };
  for (int s = 0; s < 90; ++s) {
    Cloud c;
    for (int i = 0; i < 3000; ++i) {
      const float a = ang(rng), r = rad(rng);
      c.push_back(pt(r * std::cos(a), r * std::sin(a), hgt(rng)));
    }
    if (s == 89) c = rotate(scans[7], 18.0 * M_PI / 180.0);  // revisit of place 7, turned by 3 sectors
    scans.push_back(c);
    std::vector<double> d(1200);
    orc_sc_make(&c.points[0].x, (int)c.size(), 32, d.data());
    std::vector<double> mine = sc.makeScancontext(c);
    bool same = true;
    for (int i = 0; i < 1200; ++i) same = same && (float)d[i] == (float)mine[i];
    if (s % 30 == 0 || s == 89) EXPECT(same);
    sc.makeAndSaveScancontextAndKeys(c);
    db.insert(db.end(), d.begin(), d.end());
    if (s == 20) EXPECT(sc.detectLoopClosureID().first == -1);  // fewer than NUM_EXCLUDE_RECENT + 1 entries
  }
  sc.tree_making_period_conter = 0;  // force the candidate window to refresh, as a tree rebuild would
  std::pair<int, float> hit = sc.detectLoopClosureID();
  double dist;
  int32_t id, shift;
  std::vector<double> dbf(db.size());
  for (size_t i = 0; i < db.size(); ++i) dbf[i] = (double)(float)db[i];  // the database stores float descriptors
  orc_sc_topk(dbf.data(), 90 - 50, &dbf[89 * 1200], 1, &dist, &id, &shift);
  EXPECT(id == 7 && hit.first == 7);
  EXPECT(std::fabs(sc.lastDistance() - dist) < 1e-9 && dist < 0.13);
  EXPECT(std::fabs(hit.second - (float)(shift * 6.0 * M_PI / 180.0)) < 1e-6);
  std::printf("  loop %d, distance %.4f, yaw %.3f rad\n", hit.first, sc.lastDistance(), hit.second);
}

static void test_odom_handler() {
  std::printf("OdomHandler::callback (odometry merge)\n");
  ilsm::OdomHandler h;
  auto pose = [](double yaw, double x, double y, double* p) {
    p[0] = 0, p[1] = 0, p[2] = std::sin(yaw / 2), p[3] = std::cos(yaw / 2), p[4] = x, p[5] = y, p[6] = 0;
  };
  double a[7], b[7], m[7];
  pose(0.0, 0, 0, a), pose(0.0, 0, 0, b);
  h.callback(a, b, "", m);
  EXPECT(std::fabs(m[3] - 1) < 1e-15 && m[4] == 0);
  // frame 1: intensity odometry moved 1 m forward, A-LOAM 2 m: not skipped -> intensity increment
  pose(0.0, 2, 0, a), pose(0.0, 1, 0, b);
  h.callback(a, b, "", m);
  EXPECT(std::fabs(m[4] - 1) < 1e-15);
  // frame 2: flagged "/odom_skip" -> A-LOAM increment (turn 0.1 rad, +2 m in its own frame)
  pose(0.1, 4, 0, a), pose(0.0, 1.5, 0, b);
  h.callback(a, b, "/odom_skip", m);
  EXPECT(std::fabs(m[4] - 3) < 1e-14 && std::fabs(m[2] - std::sin(0.05)) < 1e-15 && std::fabs(m[3] - std::cos(0.05)) < 1e-15);
  // frame 3: intensity again: increment expressed in the merged frame (rotated by 0.1)
  pose(0.1, 5, 0, a), pose(0.0, 2.5, 0, b);
  h.callback(a, b, "", m);
  EXPECT(std::fabs(m[4] - (3 + std::cos(0.1))) < 1e-14 && std::fabs(m[5] - std::sin(0.1)) < 1e-14);
}

static void test_messages_out() {
  std::printf("toROSMsg / odomAftMapped / aftMappedTransform (messages out)\n");
  ilsm::PointCloud<ilsm::PointXYZI> cloud;
  for (int i = 0; i < 1000; ++i) {
    ilsm::PointXYZI p;
    p.x = 0.5f * i, p.y = -1.25f * i, p.z = 3.0f + i, p.intensity = (float)(i % 64) + 0.1f;
    cloud.push_back(p);
  }
  const ilsm::PointCloud2Msg m = ilsm::toROSMsg(cloud, 12.5, "os_sensor");
  EXPECT(m.width == 1000 && m.height == 1 && m.point_step == 32 && m.row_step == 32000 && m.data.size() == 32000);
  EXPECT(m.header.frame_id == "os_sensor" && m.header.stamp == 12.5 && m.fields[3].offset == 16 && m.fields[3].datatype == 7);
  bool same = true, zeros = true;
  for (int i = 0; i < 1000; ++i) {
    float f[8];
    std::memcpy(f, &m.data[32 * (size_t)i], 32);
    same = same && f[0] == cloud[i].x && f[1] == cloud[i].y && f[2] == cloud[i].z && f[4] == cloud[i].intensity;
    zeros = zeros && f[3] == 0.f && f[5] == 0.f && f[6] == 0.f && f[7] == 0.f;
  }
  EXPECT(same && zeros);
  EXPECT(ilsm::toROSMsg(ilsm::PointCloud<ilsm::PointXYZI>(), 0.0, "map").data.empty());
  const double q[4] = {0, 0, std::sin(0.2), std::cos(0.2)}, t[3] = {1, 2, 3};
  const ilsm::OdometryMsg o = ilsm::odomAftMapped(q, t, 7.0);
  const ilsm::TransformMsg tf = ilsm::aftMappedTransform(o);
  EXPECT(o.header.frame_id == "map" && o.child_frame_id == "/aft_mapped" && o.position[2] == 3 && o.orientation_xyzw[3] == std::cos(0.2));
  EXPECT(tf.frame_id == "map" && tf.child_frame_id == "/aft_mapped" && tf.stamp == 7.0 && tf.origin[1] == 2 && tf.rotation_xyzw[2] == std::sin(0.2));
  EXPECT(ilsm::laserOdometryMsg(q, t, 1.0).header.frame_id == "camera_init");
}

static void test_config(const char* yaml_path) {
  std::printf("Config::load_yaml / load_launch (the reference's parameter keys)\n");
  ilsm::Config c;
  EXPECT(c.load_yaml(yaml_path));
  c.validate();
  EXPECT(c.image_width == 1024 && c.image_height == 64 && c.minimum_range == 0.3 && c.mapping_line_resolution == 0.4 &&
         c.mapping_plane_resolution == 0.8 && c.mapping_skip_frame == 1 && c.ground_plane_window_size == 2 &&
         c.cloud_topic == "/os_cloud_node/points");
  // overrides, comments, quoting and the launch-file form
  const char* tmp = "/tmp/ilsm_cfg_test.yaml";
  FILE* f = std::fopen(tmp, "w");
  std::fputs("intensity_feature_tracker:\n  image_width: 2048   # wide\n  cloud_topic: '/points'\n"
             "map_optimization_parameters:\n  remove_radius: 0.5\nmapping_line_resolution: 0.2\n", f);
  std::fclose(f);
  ilsm::Config d;
  EXPECT(d.load_yaml(tmp) && d.image_width == 2048 && d.cloud_topic == "/points" && d.minimum_range == 0.5 &&
         d.mapping_line_resolution == 0.2 && d.mapping_plane_resolution == 0.8);
  f = std::fopen(tmp, "w");
  std::fputs("<launch>\n  <param name=\"mapping_plane_resolution\" type=\"double\" value=\"1.6\"/>\n"
             "  <param name=\"mapping_skip_frame\" type=\"int\" value=\"2\" />\n</launch>\n", f);
  std::fclose(f);
  EXPECT(d.load_launch(tmp) && d.mapping_plane_resolution == 1.6 && d.mapping_skip_frame == 2);
  EXPECT(!d.load_yaml("/nonexistent/spot.yaml"));
  std::remove(tmp);
}

int main(int argc, char** argv) {
  try {
    if (argc > 1) test_config(argv[1]);
    test_odom_handler();
    test_messages_out();
    test_kdtree_flann();
    test_voxelgrid();
    test_ikd_tree();
    test_image_handler_and_scan_registration();
    test_scan_to_map_registration();
    test_loam_pipeline();
    test_scancontext();
  } catch (const ilsm::Error& e) {
    std::printf("ilsm::Error %d: %s\n", e.code, e.what());
    return 100;
  }
  std::printf(g_failed ? "%d check(s) FAILED\n" : "all host-mirror checks passed\n", g_failed);
  return g_failed;
}
