// slam_replay.cpp -- the full per-frame loop driven from C++ (the reference's language), no Python in the timed region:
// reads a file of organised OS0-64 frames (n_frames x 65536 x {x,y,z,intensity} float32, written by
// tools/cpp_replay.py or any other producer), uploads each frame from pinned host memory through ilsm_slam_frame and
// prints one JSON line: frames/s, the last odometry / mapped poses (bit-comparable with the Python path).
//   g++ -std=c++14 -O2 -Iinclude tools/cpp/slam_replay.cpp -o slam_replay -L<pkg> -lilsm_cuda -Wl,-rpath,<pkg>
//   ./slam_replay frames.bin [n_sequences] [pipelined]
//       n_sequences > 1: that many independent replays on concurrent threads
//       pipelined = 1  : ilsm_slam_create_async / ilsm_slam_frame_async (laserMapping as its own stage)
#include <atomic>
#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <thread>
#include <vector>

#include "ilsm.h"

#define CHECK(call)                                                             \
  do {                                                                          \
    int rc_ = (call);                                                           \
    if (rc_ != ILSM_OK) {                                                       \
      std::fprintf(stderr, "%s: %d %s\n", #call, rc_, ilsm_last_error());       \
      std::exit(2);                                                             \
    }                                                                           \
  } while (0)

struct Result {
  double q_odom[4], t_odom[3], q_map[4], t_map[3];
  double seconds;
  double phases[8];  // ilsm_slam_host_phases of the timed pass
};

static std::atomic<int> g_ready{0};

static int g_pipelined = 0;

static void replay(const float* frames, int n_frames, int pts, int warmup, int n_seq, Result* out) {
  ilsm_ctx* ctx = nullptr;
  CHECK(ilsm_create(0, &ctx));
  for (int pass = 0; pass < 2; ++pass) {  // pass 0: warm-up (allocations, module load) on the first frames
    ilsm_slam* slam = nullptr;
    if (g_pipelined) CHECK(ilsm_slam_create_async(ctx, 0.4f, 0.8f, 0.3f, 8192, &slam));
    else CHECK(ilsm_slam_create(ctx, 0.4f, 0.8f, 0.3f, 8192, &slam));
    const int n = pass == 0 ? (warmup < n_frames ? warmup : n_frames) : n_frames;
    CHECK(ilsm_sync(ctx));
    if (pass == 1) {  // every sequence has its 2.5 GB of cube slabs before anyone starts the clock (cudaMalloc stalls the device)
      g_ready.fetch_add(1);
      while (g_ready.load() < n_seq) std::this_thread::yield();
    }
    double scratch[8];
    CHECK(ilsm_slam_host_phases(slam, scratch));  // reset
    const auto t0 = std::chrono::steady_clock::now();
    for (int k = 0; k < n; ++k) {
      ilsm_slam_stats st;
      if (g_pipelined) {
        int have_prev = 0;  // the mapped pose that comes back belongs to frame k - 1
        CHECK(ilsm_slam_frame_async(slam, frames + (size_t)k * pts * 4, pts, 16, 1, out->q_odom, out->t_odom, out->q_map, out->t_map,
                                    &have_prev, &st));
      } else {
        CHECK(ilsm_slam_frame(slam, frames + (size_t)k * pts * 4, pts, 16, 1, out->q_odom, out->t_odom, out->q_map, out->t_map, &st));
      }
    }
    if (g_pipelined) {
      int have = 0;
      ilsm_slam_stats st;
      CHECK(ilsm_slam_flush(slam, out->q_map, out->t_map, &have, &st));
    }
    out->seconds = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
    CHECK(ilsm_slam_host_phases(slam, out->phases));
    ilsm_slam_destroy(slam);
  }
  ilsm_destroy(ctx);
}

int main(int argc, char** argv) {
  if (argc < 2) {
    std::fprintf(stderr, "usage: %s frames.bin [n_sequences]\n", argv[0]);
    return 1;
  }
  const int n_seq = argc > 2 ? std::atoi(argv[2]) : 1, pts = 64 * 1024;
  g_pipelined = argc > 3 ? std::atoi(argv[3]) : 0;
  FILE* f = std::fopen(argv[1], "rb");
  if (!f) return 1;
  std::fseek(f, 0, SEEK_END);
  const long bytes = std::ftell(f);
  std::fseek(f, 0, SEEK_SET);
  const int n_frames = (int)(bytes / ((long)pts * 16));
  std::vector<float> buf((size_t)n_frames * pts * 4);
  if (std::fread(buf.data(), 1, (size_t)n_frames * pts * 16, f) != (size_t)n_frames * pts * 16) return 1;
  std::fclose(f);
  CHECK(ilsm_host_register(buf.data(), buf.size() * sizeof(float)));  // frames are uploaded by DMA, as from a pinned ROS buffer
  std::vector<Result> res(n_seq);
  std::vector<std::thread> th;
  const auto t0 = std::chrono::steady_clock::now();
  for (int i = 0; i < n_seq; ++i) th.emplace_back(replay, buf.data(), n_frames, pts, 5, n_seq, &res[i]);
  for (auto& t : th) t.join();
  (void)t0;
  double slowest = 0;
  for (auto& r : res) slowest = r.seconds > slowest ? r.seconds : slowest;
  CHECK(ilsm_host_unregister(buf.data()));
  const Result& r = res[0];
  std::printf("{\"frames\": %d, \"sequences\": %d, \"pipelined\": %d, \"frames_per_s\": %.3f, \"ms_per_frame\": %.6f, "
              "\"q_map\": [%.17g, %.17g, %.17g, %.17g], \"t_map\": [%.17g, %.17g, %.17g], "
              "\"q_odom\": [%.17g, %.17g, %.17g, %.17g], \"t_odom\": [%.17g, %.17g, %.17g]}\n",
              n_frames, n_seq, g_pipelined, n_seq * n_frames / slowest, 1e3 * slowest / n_frames, r.q_map[0], r.q_map[1], r.q_map[2], r.q_map[3],
              r.t_map[0], r.t_map[1], r.t_map[2], r.q_odom[0], r.q_odom[1], r.q_odom[2], r.q_odom[3], r.t_odom[0], r.t_odom[1],
              r.t_odom[2]);
  std::fprintf(stderr, "host phases, us per frame: fe_launch %.1f  wait_prev_mapping %.1f  wait_front_end %.1f  odom_launch %.1f  "
               "wait_odom %.1f  trees+mapping %.1f  | mapping thread: launch %.1f  wait %.1f\n",
               1e6 * r.phases[0] / n_frames, 1e6 * r.phases[1] / n_frames, 1e6 * r.phases[2] / n_frames, 1e6 * r.phases[3] / n_frames,
               1e6 * r.phases[4] / n_frames, 1e6 * r.phases[5] / n_frames, 1e6 * r.phases[6] / n_frames, 1e6 * r.phases[7] / n_frames);
  return 0;
}
