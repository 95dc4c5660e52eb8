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

#include <chrono>
#include <condition_variable>
#include <new>
#include <thread>

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
  // asynchronous mapping (ilsm_slam_create_async): the cube map lives on a second, owned context; the mapping of frame k
  // runs there while the front end and the odometry of frame k + 1 run on the caller's context
  ilsm_ctx* ctx2 = nullptr;
  bool async = false, map_in_flight = false;
  DevBuf<float4> lsharp2[2], lflat2[2];  // per-frame copies of the mapping stage's inputs (frame parity)
  cudaEvent_t ev_fe[2] = {nullptr, nullptr};
  // The mapping stage has its own host thread, like the reference's laserMapping node has its own process: its ~25
  // launches per frame (3-4 us of host time each) are issued while the caller's thread launches the next front end.
  struct MapJob {
    int slot = 0, n_lsharp = 0, n_lflat = 0;
    double q_odom[4] = {0, 0, 0, 1}, t_odom[3] = {0, 0, 0};
  } job;
  struct MapResult {
    int rc = 0;
    char err[256] = "";
    double q[4] = {0, 0, 0, 1}, t[3] = {0, 0, 0};
    ilsm_reg_report report;
    ilsm_cubemap_stats cstats;
  } result;
  std::thread worker;
  std::mutex wmu;
  std::condition_variable wcv;
  bool job_posted = false, result_ready = false, stop = false;
  // host-side phase clock (profiling aid, read with ilsm_slam_host_phases): seconds accumulated per phase
  double phase_s[8] = {0, 0, 0, 0, 0, 0, 0, 0};
};

struct PhaseClock {
  double* acc;
  std::chrono::steady_clock::time_point t;
  explicit PhaseClock(double* a) : acc(a), t(std::chrono::steady_clock::now()) {}
  void lap(int k) {
    const auto n = std::chrono::steady_clock::now();
    acc[k] += std::chrono::duration<double>(n - t).count();
    t = n;
  }
};

// body of the mapping stage's thread: one process() iteration per posted job, enqueue + collect on the second context
static void mapping_worker(SlamH* sp) {
  SlamH& s = *sp;
  Ctx& c2 = s.ctx2->c;
  cudaSetDevice(c2.device);
  for (;;) {
    SlamH::MapJob j;
    {
      std::unique_lock<std::mutex> lk(s.wmu);
      s.wcv.wait(lk, [&] { return s.job_posted || s.stop; });
      if (s.stop) return;
      j = s.job;
      s.job_posted = false;
    }
    SlamH::MapResult r;
    {
      std::lock_guard<std::mutex> lk2(c2.mu);
      ilsm_reg_opts mo;
      ilsm_reg_opts_default(&mo);
      memset(&r.report, 0, sizeof(r.report)), memset(&r.cstats, 0, sizeof(r.cstats));
      cudaError_t e = cudaStreamWaitEvent(c2.stream, s.ev_fe[j.slot], 0);
      r.rc = e == cudaSuccess ? ILSM_OK : fail_cuda(e, "cudaStreamWaitEvent(mapping stage)");
      PhaseClock wclk(s.phase_s);  // [6] mapping-stage launches, [7] wait for the mapped pose (written by this thread only)
      if (!r.rc)
        r.rc = cubemap_frame_enqueue(s.cube->m, reinterpret_cast<const float*>(s.lsharp2[j.slot].p), j.n_lsharp,
                                     reinterpret_cast<const float*>(s.lflat2[j.slot].p), j.n_lflat, 16, j.q_odom, j.t_odom, mo, true, true,
                                     s.ctx->ev_join);  // the stacks come from the caller's side stream (under its odometry)
      wclk.lap(6);
      if (!r.rc) r.rc = cubemap_frame_collect(s.cube->m, r.q, r.t, &r.report, &r.cstats);
      wclk.lap(7);
      if (r.rc) snprintf(r.err, sizeof(r.err), "%s", ilsm_last_error());  // the error text is thread-local: carry it over
    }
    {
      std::lock_guard<std::mutex> lk(s.wmu);
      s.result = r;
      s.result_ready = true;
    }
    s.wcv.notify_all();
  }
}

// wait for the mapping stage's answer to the job in flight and hand it out
static int take_mapping_result(SlamH& s, double q_map[4], double t_map[3], ilsm_slam_stats* stats) {
  std::unique_lock<std::mutex> lk(s.wmu);
  s.wcv.wait(lk, [&] { return s.result_ready; });
  s.result_ready = false;
  s.map_in_flight = false;
  const SlamH::MapResult& r = s.result;
  if (r.rc) return fail(r.rc, r.err);
  for (int i = 0; i < 4; ++i) q_map[i] = r.q[i];
  for (int i = 0; i < 3; ++i) t_map[i] = r.t[i];
  if (stats) stats->mapping = r.report, stats->cubemap = r.cstats;
  return ILSM_OK;
}

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

ILSM_API int ilsm_slam_create_async(ilsm_ctx* ctx, float line_res, float plane_res, float min_range, int cube_capacity,
                                    ilsm_slam** out) {
  ilsm_slam* h = nullptr;
  int rc = slam_create_common(ctx, min_range, out, &h);
  if (rc) return rc;
  h->s.async = true;
  if ((rc = ilsm_create(ctx->c.device, &h->s.ctx2)) ||
      (rc = ilsm_cubemap_create(h->s.ctx2, line_res, plane_res, cube_capacity, &h->s.cube))) {
    ilsm_slam_destroy(h);
    return rc;
  }
  for (int i = 0; i < 2; ++i)
    if (cudaEventCreateWithFlags(&h->s.ev_fe[i], cudaEventDisableTiming) != cudaSuccess) {
      ilsm_slam_destroy(h);
      return fail(ILSM_ERR_CUDA, "slam_create_async: event creation failed");
    }
  h->s.worker = std::thread(mapping_worker, &h->s);
  *out = h;
  return ILSM_OK;
}

ILSM_API void ilsm_slam_destroy(ilsm_slam* slam) {
  if (!slam) return;
  SlamH& s = slam->s;
  if (s.worker.joinable()) {
    {
      std::unique_lock<std::mutex> lk(s.wmu);
      s.wcv.wait(lk, [&] { return !s.job_posted; });  // a posted job is taken (and finished) before the thread is told to stop
      s.stop = true;
    }
    s.wcv.notify_all();
    s.worker.join();
  }
  if (s.ctx2) {  // let the mapping stage drain before anything it reads goes away
    std::lock_guard<std::mutex> lk2(s.ctx2->c.mu);
    cudaSetDevice(s.ctx2->c.device);
    cudaStreamSynchronize(s.ctx2->c.stream);
    cudaStreamSynchronize(s.ctx2->c.aux);
  }
  {
    std::lock_guard<std::mutex> lk(s.ctx->mu);
    cudaSetDevice(s.ctx->device);
    cudaStreamSynchronize(s.ctx->stream);
    s.last_corner.release(), s.last_surf.release();
    s.sharp.release(), s.flat.release(), s.lsharp.release();
    for (int i = 0; i < 2; ++i) {
      s.lsharp2[i].release(), s.lflat2[i].release();
      if (s.ev_fe[i]) cudaEventDestroy(s.ev_fe[i]);
    }
  }
  if (s.cube) ilsm_cubemap_destroy(s.cube);
  if (s.mapopt) ilsm_mapopt_destroy(s.mapopt);
  if (s.ctx2) ilsm_destroy(s.ctx2);
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
                           double q_odom[4], double t_odom[3], double q_map[4], double t_map[3], ilsm_slam_stats* stats,
                           int* have_prev = nullptr);
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

ILSM_API int ilsm_slam_frame_async(ilsm_slam* slam, const float* xyzi, int n, int stride_bytes, int use_aloam, double q_odom[4],
                                   double t_odom[3], double q_map_prev[4], double t_map_prev[3], int* have_prev,
                                   ilsm_slam_stats* stats) {
  if (n < 0 || stride_bytes < 12 || stride_bytes % 4) return fail(ILSM_ERR_INVALID_ARG, "slam_frame_async: bad n/stride");
  if (!have_prev) return fail(ILSM_ERR_INVALID_ARG, "slam_frame_async: null have_prev");
  return slam_frame_impl(slam, xyzi, n, stride_bytes, nullptr, use_aloam, q_odom, t_odom, q_map_prev, t_map_prev, stats, have_prev);
}

ILSM_API int ilsm_slam_flush(ilsm_slam* slam, double q_map[4], double t_map[3], int* have, ilsm_slam_stats* stats) {
  if (!slam || !q_map || !t_map || !have) return fail(ILSM_ERR_INVALID_ARG, "slam_flush: null argument");
  SlamH& s = slam->s;
  *have = 0;
  if (!s.async || !s.map_in_flight) return ILSM_OK;
  std::lock_guard<std::mutex> lk(s.ctx->mu);
  int rc = take_mapping_result(s, q_map, t_map, stats);
  if (rc) return rc;
  *have = 1;
  return ILSM_OK;
}

}  // extern "C"

static int slam_frame_impl(ilsm_slam* slam, const float* xyzi, int n, int stride_bytes, const ilsm_pc2_layout* pc2, int use_aloam,
                           double q_odom[4], double t_odom[3], double q_map[4], double t_map[3], ilsm_slam_stats* stats,
                           int* have_prev) {
  if (!slam || (n > 0 && !xyzi) || !q_odom || !t_odom || !q_map || !t_map)
    return fail(ILSM_ERR_INVALID_ARG, "slam_frame: null argument");
  SlamH& s = slam->s;
  if (s.async != (have_prev != nullptr))
    return fail(ILSM_ERR_STATE, "slam_frame: ilsm_slam_create_async handles take ilsm_slam_frame_async (and only those)");
  if (have_prev) *have_prev = 0;
  Ctx& c = *s.ctx;
  std::lock_guard<std::mutex> lk(c.mu);
  ILSM_CUDA(cudaSetDevice(c.device));
  if (stats) memset(stats, 0, sizeof(*stats));
  int rc;
  PhaseClock clk(s.phase_s);
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
  clk.lap(0);  // upload + front-end launches
  int* pin = reinterpret_cast<int*>(c.pinned.p);
  ILSM_CUDA(cudaMemcpyAsync(pin, c.fe.counts.p, 8 * sizeof(int), cudaMemcpyDeviceToHost, c.stream));
  ILSM_CUDA(cudaStreamSynchronize(c.stream));
  clk.lap(2);  // wait for the front end (feature counts)
  if (s.async && s.map_in_flight) {
    // the previous frame's mapped pose: the mapping stage's thread has had this frame's upload and front end to finish
    if ((rc = take_mapping_result(s, q_map, t_map, stats))) return rc;
    *have_prev = 1;
  }
  clk.lap(1);  // (pipelined) wait for the previous frame's mapping
  const int n_sharp = pin[1], n_lsharp = pin[2], n_flat = pin[3], n_lflat = pin[4];
  if (pin[5]) return fail(ILSM_ERR_INVALID_ARG, "slam_frame: a ring segment exceeds the supported size");
  if (stats) {
    stats->n_cloud = pin[0], stats->n_sharp = n_sharp, stats->n_less_sharp = n_lsharp, stats->n_flat = n_flat;
    stats->n_less_flat = n_lflat;
  }
  // the mapping stage's inputs: in asynchronous mode they are per-frame copies (frame parity), because the mapping of this
  // frame still reads them while the next frame's front end runs
  const int slot = (int)(s.frames & 1);
  DevBuf<float4>& lsharp_buf = s.async ? s.lsharp2[slot] : s.lsharp;
  if ((rc = s.sharp.reserve(n_sharp + 4)) || (rc = s.flat.reserve(n_flat + 4)) || (rc = lsharp_buf.reserve(n_lsharp + 4))) return rc;
  if (s.async && (rc = s.lflat2[slot].reserve(n_lflat + 4))) return rc;
  float4* const d_lsharp = lsharp_buf.p;
  if ((rc = c.gather_dev(c.fe.cloud.p, c.fe.sharp.p, c.fe.counts.p, 1, n_sharp, s.sharp.p)) ||
      (rc = c.gather_dev(c.fe.cloud.p, c.fe.lsharp.p, c.fe.counts.p, 2, n_lsharp, d_lsharp)) ||
      (rc = c.gather_dev(c.fe.cloud.p, c.fe.flat.p, c.fe.counts.p, 3, n_flat, s.flat.p)))
    return rc;
  if (s.async) {
    if (n_lflat) ILSM_CUDA(cudaMemcpyAsync(s.lflat2[slot].p, c.fe.lflat.p, (size_t)n_lflat * sizeof(float4), cudaMemcpyDeviceToDevice, c.stream));
    ILSM_CUDA(cudaEventRecord(s.ev_fe[slot], c.stream));
  }
  // ---- the mapping stacks (VoxelGrid of the less-sharp / less-flat clouds, laserMapping.cpp:608-616) depend only on the
  // front end: they run on the side stream while the odometry solves on the main one.  Clouds above 16384 points take
  // the tiled multi-block VoxelGrid (no size limit below 2^24 points, like the reference).
  CubeMapH* cmp = s.cube ? &s.cube->m : nullptr;
  if (cmp) {
    CubeMapH& cm = *cmp;  // (pipelined mode: the mapping thread is idle here -- its previous result has been taken above)
    if ((rc = cm.stack_c.reserve(n_lsharp + 4)) || (rc = cm.stack_s.reserve(n_lflat + 4))) return rc;
    ILSM_CUDA(cudaEventRecord(c.ev_fork, c.stream));
    ILSM_CUDA(cudaStreamWaitEvent(c.aux, c.ev_fork, 0));
    // pipelined mode: the previous frame's deferred insertion (second context's side stream) still reads the stacks
    if (s.async && cm.tail_pending) ILSM_CUDA(cudaStreamWaitEvent(c.aux, cm.ev_tail, 0));
    ILSM_CUDA(cudaMemsetAsync(cm.stack_n.p, 0, 4 * sizeof(int), c.aux));
    // Pipelined mode: the side stream reads the PER-FRAME copy of the less-flat cloud -- the next call's front end
    // rewrites fe.lflat (and resets the front end's error word) while this VoxelGrid may still be running, and nothing
    // orders the main stream after it there.  The VoxelGrid reports into the cube map's own error word.
    const float4* d_lflat_vg = s.async ? s.lflat2[slot].p : c.fe.lflat.p;
    if ((rc = c.voxelgrid_pair_dev(reinterpret_cast<const float*>(d_lsharp), n_lsharp, cm.line_res, cm.stack_c.p,
                                   reinterpret_cast<const float*>(d_lflat_vg), n_lflat, cm.plane_res, cm.stack_s.p, 16, 3,
                                   cm.stack_n.p, c.aux, cm.err.p)))
      return rc;
    ILSM_CUDA(cudaEventRecord(c.ev_join, c.aux));
  }
  // ---- laserOdometry
  // the previous frame's deferred map insertion (side stream) reads the mapped pose from the LM state the odometry is
  // about to overwrite: order the main stream after it (long finished by now -- it overlapped this frame's front end)
  if (cmp && !s.async && cmp->tail_pending) ILSM_CUDA(cudaStreamWaitEvent(c.stream, cmp->ev_tail, 0));
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
      clk.lap(3);  // gathers, stack VoxelGrid, odometry launches
      ILSM_CUDA(cudaStreamSynchronize(c.stream));
      clk.lap(4);  // wait for the odometry
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
  if ((rc = build_pair_dev(&s.last_corner, reinterpret_cast<const float*>(d_lsharp), n_lsharp, &s.last_surf,
                           reinterpret_cast<const float*>(c.fe.lflat.p), n_lflat, 16, kOdomCell)))
    return rc;
  // ---- mapping (mapping_skip_frame = 1: every frame is published, laserOdometry.cpp:810-833)
  if (s.async) {
    // laserMapping as its own stage (the reference runs it as its own node): hand this frame to the mapping thread and
    // return; its mapped pose is collected by the next call (or by ilsm_slam_flush)
    {
      std::lock_guard<std::mutex> lkw(s.wmu);
      s.job.slot = slot, s.job.n_lsharp = n_lsharp, s.job.n_lflat = n_lflat;
      for (int i = 0; i < 4; ++i) s.job.q_odom[i] = q_odom[i];
      for (int i = 0; i < 3; ++i) s.job.t_odom[i] = t_odom[i];
      s.job_posted = true;
      s.map_in_flight = true;
    }
    s.wcv.notify_all();
  } else if (cmp) {  // laserMapping: rolling cube map
    ilsm_reg_opts mo;
    ilsm_reg_opts_default(&mo);
    rc = cubemap_frame_core(*cmp, reinterpret_cast<const float*>(d_lsharp), n_lsharp,
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
  clk.lap(5);  // tree builds + mapping stage (synchronous: enqueue and wait; pipelined: enqueue only)
  if (rc) return rc;
  s.frames++;
  return ILSM_OK;
}

extern "C" ILSM_API int ilsm_slam_host_phases(ilsm_slam* slam, double out8[8]) {
  if (!slam || !out8) return fail(ILSM_ERR_INVALID_ARG, "slam_host_phases: null argument");
  for (int i = 0; i < 8; ++i) out8[i] = slam->s.phase_s[i], slam->s.phase_s[i] = 0;
  return ILSM_OK;
}
