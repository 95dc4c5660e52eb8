// pipeline.cu -- the full per-frame loop of the reference's three LOAM nodes chained on one GPU (BASELINE config 2):
//   scanRegistration  laserCloudHandler          scanRegistration.cpp:189-669   -> Ctx::features_dev
//   laserOdometry     main loop                  laserOdometry.cpp:256-845      -> Ctx::odometry_dev + pose composition
//   laserMapping      process()                  laserMapping.cpp:233-1166      -> cubemap_frame_core
// The topics between the nodes (/laser_cloud_sharp, /laser_cloud_less_sharp, /laser_cloud_flat, /laser_cloud_less_flat,
// /laser_cloud_corner_last, /laser_cloud_surf_last, /laser_odom_to_init) become device buffers: the organised frame is
// the only upload, the two poses and a few counters the only downloads.  Host synchronisation points per frame: the
// feature counts (sizes of the next launches), the odometry pose (the host owns q_w_curr / t_w_curr and the cube
// window indices, as the reference nodes do) and the mapped pose.
#include <string.h>

#include <new>

#include "ilsm_cubemap.hpp"

struct ilsm_mapopt;

namespace ilsm {

int mapopt_frame_core(ilsm_mapopt* mo, const float* frame_xyz, int n, int stride_bytes, const float* plane_xyz, int n_plane,
                      int plane_stride_bytes, bool from_host, const double q_wodom[4], const double t_wodom[3], double q_w[4],
                      double t_w[3], const ilsm_ground_opts* gopts, ilsm_mapopt_stats* stats);  // mapopt.cu

// voxel edge of the "last frame" search structures: the odometry gate is 5 m (DISTANCE_SQ_THRESHOLD 25,
// laserOdometry.cpp:31), the clouds are sparse (a few thousand points)
constexpr float kOdomCell = 1.0f;

struct SlamH {
  Ctx* ctx = nullptr;
  ilsm_cubemap* cube = nullptr;   // mapping = laserMapping (rolling cube map)
  ilsm_mapopt* mapopt = nullptr;  // mapping = mapOptimization (ground map), the node spot.launch starts
  Map last_corner, last_surf;  // kdtreeCornerLast / kdtreeSurfLast (laserOdometry.cpp:807-808)
  DevBuf<float4> sharp, flat, lsharp;
  bool inited = false;        // systemInited (laserOdometry.cpp:384)
  double para_q[4] = {0, 0, 0, 1}, para_t[3] = {0, 0, 0};  // q_last_curr / t_last_curr, kept between frames (:51-52)
  QuatH q_w_curr{0, 0, 0, 1};
  double t_w_curr[3] = {0, 0, 0};
  float min_range = 0.3f;
  long long frames = 0;
};

}  // namespace ilsm

using namespace ilsm;

struct ilsm_slam {
  SlamH s;
};

extern "C" {

static int slam_create_common(ilsm_ctx* ctx, float min_range, ilsm_slam** out, ilsm_slam** h_out) {
  if (!ctx || !out) return fail(ILSM_ERR_INVALID_ARG, "slam_create: null argument");
  ilsm_slam* h = new (std::nothrow) ilsm_slam();
  if (!h) return fail(ILSM_ERR_OUT_OF_MEMORY, "host allocation failed");
  h->s.ctx = &ctx->c;
  h->s.min_range = min_range > 0.f ? min_range : 0.3f;
  int rc;
  {
    std::lock_guard<std::mutex> lk(ctx->c.mu);
    cudaSetDevice(ctx->c.device);
    if (!(rc = h->s.last_corner.init(&ctx->c))) rc = h->s.last_surf.init(&ctx->c);
  }
  if (rc) {
    delete h;
    return rc;
  }
  *h_out = h;
  return ILSM_OK;
}

ILSM_API int ilsm_slam_create(ilsm_ctx* ctx, float line_res, float plane_res, float min_range, int cube_capacity,
                              ilsm_slam** out) {
  ilsm_slam* h = nullptr;
  int rc = slam_create_common(ctx, min_range, out, &h);
  if (rc) return rc;
  if ((rc = ilsm_cubemap_create(ctx, line_res, plane_res, cube_capacity, &h->s.cube))) {
    ilsm_slam_destroy(h);
    return rc;
  }
  *out = h;
  return ILSM_OK;
}

ILSM_API int ilsm_slam_create_mapopt(ilsm_ctx* ctx, float voxel_leaf, float downsample_size, float min_range, ilsm_slam** out) {
  ilsm_slam* h = nullptr;
  int rc = slam_create_common(ctx, min_range, out, &h);
  if (rc) return rc;
  if ((rc = ilsm_mapopt_create(ctx, voxel_leaf, downsample_size, &h->s.mapopt))) {
    ilsm_slam_destroy(h);
    return rc;
  }
  *out = h;
  return ILSM_OK;
}

ILSM_API void ilsm_slam_destroy(ilsm_slam* slam) {
  if (!slam) return;
  SlamH& s = slam->s;
  {
    std::lock_guard<std::mutex> lk(s.ctx->mu);
    cudaSetDevice(s.ctx->device);
    cudaStreamSynchronize(s.ctx->stream);
    s.last_corner.release(), s.last_surf.release();
    s.sharp.release(), s.flat.release(), s.lsharp.release();
  }
  if (s.cube) ilsm_cubemap_destroy(s.cube);
  if (s.mapopt) ilsm_mapopt_destroy(s.mapopt);
  delete slam;
}

ILSM_API ilsm_cubemap* ilsm_slam_cubemap(ilsm_slam* slam) { return slam ? slam->s.cube : nullptr; }
ILSM_API ilsm_mapopt* ilsm_slam_mapopt(ilsm_slam* slam) { return slam ? slam->s.mapopt : nullptr; }

}  // extern "C"
namespace ilsm {
int check_pc2_layout(const ilsm_pc2_layout* l);  // api.cu
}
// pc2 != nullptr: `xyzi` is a sensor_msgs/PointCloud2 data blob, unpacked on the device into packed points
static int slam_frame_impl(ilsm_slam* slam, const float* xyzi, int n, int stride_bytes, const ilsm_pc2_layout* pc2, int use_aloam,
                           double q_odom[4], double t_odom[3], double q_map[4], double t_map[3], ilsm_slam_stats* stats);
extern "C" {

ILSM_API int ilsm_slam_frame(ilsm_slam* slam, const float* xyzi, int n, int stride_bytes, int use_aloam, double q_odom[4],
                             double t_odom[3], double q_map[4], double t_map[3], ilsm_slam_stats* stats) {
  if (n < 0 || stride_bytes < 12 || stride_bytes % 4) return fail(ILSM_ERR_INVALID_ARG, "slam_frame: bad n/stride");
  return slam_frame_impl(slam, xyzi, n, stride_bytes, nullptr, use_aloam, q_odom, t_odom, q_map, t_map, stats);
}

ILSM_API int ilsm_slam_frame_pc2(ilsm_slam* slam, const uint8_t* data, int n_points, const ilsm_pc2_layout* layout, int use_aloam,
                                 double q_odom[4], double t_odom[3], double q_map[4], double t_map[3], ilsm_slam_stats* stats) {
  if (n_points < 0) return fail(ILSM_ERR_INVALID_ARG, "slam_frame_pc2: bad n_points");
  int rc = check_pc2_layout(layout);
  if (rc) return rc;
  return slam_frame_impl(slam, reinterpret_cast<const float*>(data), n_points, layout->point_step, layout, use_aloam, q_odom, t_odom,
                         q_map, t_map, stats);
}

}  // extern "C"

static int slam_frame_impl(ilsm_slam* slam, const float* xyzi, int n, int stride_bytes, const ilsm_pc2_layout* pc2, int use_aloam,
                           double q_odom[4], double t_odom[3], double q_map[4], double t_map[3], ilsm_slam_stats* stats) {
  if (!slam || (n > 0 && !xyzi) || !q_odom || !t_odom || !q_map || !t_map)
    return fail(ILSM_ERR_INVALID_ARG, "slam_frame: null argument");
  SlamH& s = slam->s;
  Ctx& c = *s.ctx;
  std::lock_guard<std::mutex> lk(c.mu);
  ILSM_CUDA(cudaSetDevice(c.device));
  if (stats) memset(stats, 0, sizeof(*stats));
  int rc;
  // ---- scanRegistration: the frame is the only upload
  const size_t bytes = (size_t)n * stride_bytes;
  if ((rc = c.fe.raw.reserve((pc2 ? (size_t)n * 4 : bytes / 4) + 4))) return rc;
  if (pc2 && (rc = c.fe.pc2.reserve(bytes + 16))) return rc;
  // the previous frame's tree builds (on the maps' own streams, overlapping its mapping step) read lsharp / fe.lflat:
  // order this frame's front end after them
  if ((rc = s.last_corner.wait_ready(c.stream)) || (rc = s.last_surf.wait_ready(c.stream))) return rc;
  if (pc2) {  // the message blob as it is; pcl::fromROSMsg's repacking (scanRegistration.cpp:235) runs as a kernel
    if (bytes) ILSM_CUDA(cudaMemcpyAsync(c.fe.pc2.p, xyzi, bytes, cudaMemcpyHostToDevice, c.stream));
    if ((rc = c.pc2_unpack_dev(c.fe.pc2.p, n, *pc2, reinterpret_cast<float4*>(c.fe.raw.p)))) return rc;
    stride_bytes = 16;
  } else if (bytes) {
    ILSM_CUDA(cudaMemcpyAsync(c.fe.raw.p, xyzi, bytes, cudaMemcpyHostToDevice, c.stream));
  }
  if ((rc = c.features_dev(c.fe.raw.p, n, stride_bytes, s.min_range))) return rc;
  int* pin = reinterpret_cast<int*>(c.pinned.p);
  ILSM_CUDA(cudaMemcpyAsync(pin, c.fe.counts.p, 8 * sizeof(int), cudaMemcpyDeviceToHost, c.stream));
  ILSM_CUDA(cudaStreamSynchronize(c.stream));
  const int n_sharp = pin[1], n_lsharp = pin[2], n_flat = pin[3], n_lflat = pin[4];
  if (pin[5]) return fail(ILSM_ERR_INVALID_ARG, "slam_frame: a ring segment exceeds the supported size");
  if (stats) {
    stats->n_cloud = pin[0], stats->n_sharp = n_sharp, stats->n_less_sharp = n_lsharp, stats->n_flat = n_flat;
    stats->n_less_flat = n_lflat;
  }
  if ((rc = s.sharp.reserve(n_sharp + 4)) || (rc = s.flat.reserve(n_flat + 4)) || (rc = s.lsharp.reserve(n_lsharp + 4))) return rc;
  if ((rc = c.gather_dev(c.fe.cloud.p, c.fe.sharp.p, c.fe.counts.p, 1, n_sharp, s.sharp.p)) ||
      (rc = c.gather_dev(c.fe.cloud.p, c.fe.lsharp.p, c.fe.counts.p, 2, n_lsharp, s.lsharp.p)) ||
      (rc = c.gather_dev(c.fe.cloud.p, c.fe.flat.p, c.fe.counts.p, 3, n_flat, s.flat.p)))
    return rc;
  // ---- the mapping stacks (VoxelGrid of the less-sharp / less-flat clouds, laserMapping.cpp:608-616) depend only on the
  // front end: they run on the side stream while the odometry solves on the main one
  if (n_lsharp > 16384 || n_lflat > 16384) return fail(ILSM_ERR_INVALID_ARG, "slam_frame: feature cloud exceeds 16384 points");
  CubeMapH* cmp = s.cube ? &s.cube->m : nullptr;
  if (cmp) {
    CubeMapH& cm = *cmp;
    if ((rc = cm.stack_c.reserve(n_lsharp + 4)) || (rc = cm.stack_s.reserve(n_lflat + 4))) return rc;
    ILSM_CUDA(cudaEventRecord(c.ev_fork, c.stream));
    ILSM_CUDA(cudaStreamWaitEvent(c.aux, c.ev_fork, 0));
    ILSM_CUDA(cudaMemsetAsync(cm.stack_n.p, 0, 4 * sizeof(int), c.aux));
    if ((rc = c.voxelgrid_pair_dev(reinterpret_cast<const float*>(s.lsharp.p), n_lsharp, cm.line_res, cm.stack_c.p,
                                   reinterpret_cast<const float*>(c.fe.lflat.p), n_lflat, cm.plane_res, cm.stack_s.p, 16, 3,
                                   cm.stack_n.p, c.aux)))
      return rc;
    ILSM_CUDA(cudaEventRecord(c.ev_join, c.aux));
  }
  // ---- laserOdometry
  // the previous frame's deferred map insertion (side stream) reads the mapped pose from the LM state the odometry is
  // about to overwrite: order the main stream after it (long finished by now -- it overlapped this frame's front end)
  if (cmp && cmp->tail_pending) ILSM_CUDA(cudaStreamWaitEvent(c.stream, cmp->ev_tail, 0));
  if (!s.inited) {
    s.inited = true;
  } else {
    if (use_aloam) {  // the fork optimises only on frames flagged "skip_intensity" (laserOdometry.cpp:406-417)
      ilsm_reg_opts oo;
      ilsm_reg_opts_default(&oo);  // 2 passes x max 4 iterations, Huber 0.1 (laserOdometry.cpp:417,644,705-710)
      double* pin_pose = reinterpret_cast<double*>(c.pinned.p + 2048);  // para_q / para_t -> LmState::xq, xt
      for (int i = 0; i < 4; ++i) pin_pose[i] = s.para_q[i];
      for (int i = 0; i < 3; ++i) pin_pose[4 + i] = s.para_t[i];
      ILSM_CUDA(cudaMemcpyAsync(c.lm.p->xq, pin_pose, 7 * sizeof(double), cudaMemcpyHostToDevice, c.stream));
      if ((rc = c.odometry_dev(&s.last_corner, &s.last_surf, reinterpret_cast<const float*>(s.sharp.p), n_sharp,
                               reinterpret_cast<const float*>(s.flat.p), n_flat, 16, oo)))
        return rc;
      unsigned char* pb = c.pinned.p;
      ILSM_CUDA(cudaMemcpyAsync(pb, c.lm.p->xq, 7 * sizeof(double), cudaMemcpyDeviceToHost, c.stream));
      ILSM_CUDA(cudaMemcpyAsync(pb + 64, &c.lm.p->report, sizeof(ilsm_reg_report), cudaMemcpyDeviceToHost, c.stream));
      ILSM_CUDA(cudaStreamSynchronize(c.stream));
      const double* o7 = reinterpret_cast<const double*>(pb);
      for (int i = 0; i < 4; ++i) s.para_q[i] = o7[i];
      for (int i = 0; i < 3; ++i) s.para_t[i] = o7[4 + i];
      if (stats) {
        memcpy(&stats->odometry, pb + 64, sizeof(ilsm_reg_report));
        stats->odometry.passes = oo.outer_iterations;
        stats->ran_odometry = 1;
      }
    }
    // t_w_curr = t_w_curr + q_w_curr * t_last_curr;  q_w_curr = q_w_curr * q_last_curr   (laserOdometry.cpp:716-717)
    double r[3];
    qrot_h(s.q_w_curr, s.para_t, r);
    for (int i = 0; i < 3; ++i) s.t_w_curr[i] = s.t_w_curr[i] + r[i];
    s.q_w_curr = qmul_h(s.q_w_curr, QuatH{s.para_q[0], s.para_q[1], s.para_q[2], s.para_q[3]});
  }
  q_odom[0] = s.q_w_curr.x, q_odom[1] = s.q_w_curr.y, q_odom[2] = s.q_w_curr.z, q_odom[3] = s.q_w_curr.w;
  for (int i = 0; i < 3; ++i) t_odom[i] = s.t_w_curr[i];
  // laserCloudCornerLast = cornerPointsLessSharp, laserCloudSurfLast = surfPointsLessFlat; rebuild both trees (:793-808)
  if ((rc = s.last_corner.build_dev(reinterpret_cast<const float*>(s.lsharp.p), n_lsharp, 16, kOdomCell)) ||
      (rc = s.last_surf.build_dev(reinterpret_cast<const float*>(c.fe.lflat.p), n_lflat, 16, kOdomCell)))
    return rc;
  // ---- mapping (mapping_skip_frame = 1: every frame is published, laserOdometry.cpp:810-833)
  if (cmp) {  // laserMapping: rolling cube map
    ilsm_reg_opts mo;
    ilsm_reg_opts_default(&mo);
    rc = cubemap_frame_core(*cmp, reinterpret_cast<const float*>(s.lsharp.p), n_lsharp,
                            reinterpret_cast<const float*>(c.fe.lflat.p), n_lflat, 16, q_odom, t_odom, q_map, t_map, mo,
                            stats ? &stats->mapping : nullptr, stats ? &stats->cubemap : nullptr, true, true, c.ev_join);
  } else {    // mapOptimization: ground extraction from the frame already on the device + the less-flat cloud
    ilsm_mapopt_stats ms;
    rc = mapopt_frame_core(s.mapopt, c.fe.raw.p, n, stride_bytes, reinterpret_cast<const float*>(c.fe.lflat.p), n_lflat, 16, false,
                           q_odom, t_odom, q_map, t_map, nullptr, &ms);
    if (!rc && stats) {
      stats->mapping.passes = ms.ran_optimization;
      stats->mapping.pass[0] = ms.solve;
      stats->cubemap.ran_optimization = ms.ran_optimization;
      stats->cubemap.n_map_surf = ms.map_size;
      stats->cubemap.n_stack_surf = ms.n_query;
      stats->cubemap.n_valid = ms.converged;
    }
  }
  if (rc) return rc;
  s.frames++;
  return ILSM_OK;
}
