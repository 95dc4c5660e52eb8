// ilsm_host.hpp -- host-side objects behind the opaque C handles (ilsm_ctx, ilsm_map).
#pragma once
#include <cuda_runtime.h>
#include <limits.h>
#include <stdint.h>

#include <mutex>
#include <string>

#include "../../include/ilsm.h"
#include "ilsm_internal.cuh"

namespace ilsm {

int fail(int code, const char* msg);
int fail_cuda(cudaError_t e, const char* where);
int check_launch(const char* where);
void count_launches(int k);

#define ILSM_CUDA(call)                                        \
  do {                                                         \
    cudaError_t e__ = (call);                                  \
    if (e__ != cudaSuccess) return ::ilsm::fail_cuda(e__, #call); \
  } while (0)

// Launch with programmatic stream serialization: the kernel's launch latency overlaps the tail of its predecessor
// (the short kernels of a registration are launch-latency-bound).  The kernel must start with pdl_entry().
template <typename... KArgs, typename... Args>
inline cudaError_t launch_pdl(void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t s, Args&&... args) {
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = grid, cfg.blockDim = block, cfg.dynamicSmemBytes = smem, cfg.stream = s;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr, cfg.numAttrs = 1;
  return cudaLaunchKernelEx(&cfg, kernel, static_cast<KArgs>(args)...);
}

// Grow-only device buffer (HBM is plentiful: 180 GB; reallocation would serialise the stream).
template <typename T>
struct DevBuf {
  T* p = nullptr;
  size_t cap = 0;
  int reserve(size_t n) {
    if (n <= cap) return ILSM_OK;
    size_t want = n + n / 2 + 64;
    T* np = nullptr;
    cudaError_t e = cudaMalloc(&np, want * sizeof(T));
    if (e != cudaSuccess) {
      cudaGetLastError();
      return fail(ILSM_ERR_OUT_OF_MEMORY, "cudaMalloc failed");
    }
    if (p) cudaFree(p);  // implicit device sync: only on growth
    p = np;
    cap = want;
    return ILSM_OK;
  }
  void release() {
    if (p) cudaFree(p);
    p = nullptr;
    cap = 0;
  }
};

template <typename T>
struct PinnedBuf {
  T* p = nullptr;
  size_t cap = 0;
  int reserve(size_t n) {
    if (n <= cap) return ILSM_OK;
    size_t want = n + n / 2 + 64;
    T* np = nullptr;
    if (cudaMallocHost(&np, want * sizeof(T)) != cudaSuccess) {
      cudaGetLastError();
      return fail(ILSM_ERR_OUT_OF_MEMORY, "cudaMallocHost failed");
    }
    if (p) cudaFreeHost(p);
    p = np;
    cap = want;
    return ILSM_OK;
  }
  void release() {
    if (p) cudaFreeHost(p);
    p = nullptr;
    cap = 0;
  }
};

// Levenberg-Marquardt state that lives in HBM for the duration of a registration: the trust-region loop of
// ceres::Solve runs on the device, one evaluation kernel per iteration, no host round trip.
struct LmState {
  double xq[4], xt[3];  // accepted pose (q = x,y,z,w)
  double cq[4], ct[3];  // candidate pose: what the next evaluation kernel evaluates
  double cost;          // cost at x
  double H[21], g[6];   // unscaled J^T J (upper triangle, row-major) and J^T r at x
  double scale[6];      // Jacobi scaling fixed at iteration 0
  double diag[6];       // LM diagonal (of the scaled J^T J) kept while reuse_diagonal
  double radius, decrease_factor, model_cost_change;
  double huber_a;
  double initial_cost;
  int status;  // 0 running, 1 + ilsm_termination when finished
  int phase;   // 0: next evaluation is the initial one, 1: trial step
  int iteration, max_iter, invalid_run, reuse_diag;
  int n_success, n_unsuccess, n_evals;
  int n_edge, n_plane;
  int pass;  // which report slot the running solve fills
  unsigned ticket;
  int pad;
  ilsm_reg_report report;
  long long dbg[64];  // clock64() phase stamps written when ILSM_DEBUG_TIMING is compiled in (profiling aid)
};

// Where the association kernel of the FIRST pass takes the pose from, and where the solve kernel of the LAST pass
// leaves the result: folding these into the two kernels removes the 1-warp pose upload / download launches.
struct PoseSrc {
  const double* dptr = nullptr;  // mode 2: 7 doubles on the device
  double v[7] = {0, 0, 0, 1, 0, 0, 0};  // mode 1: by value (host-pointer entry points)
  int mode = 0;                  // 0: the accepted pose already in the LM state
};
struct PoseDst {
  double* d_pose7 = nullptr;           // pose after the last pass (nullptr: stays in the LM state only)
  ilsm_reg_report* d_report = nullptr;  // the whole report (nullptr: stays in the LM state only)
};

struct Ctx;
struct Map;
int build_pair_dev(Map* a, const float* d_a, int na, Map* b, const float* d_b, int nb, int stride_bytes, float cell_size);

struct Map {
  Ctx* ctx = nullptr;
  int n = 0;
  float cell = 1.f, inv_cell = 1.f;
  uint32_t table_size = 0;
  int log2_size = 0;
  DevBuf<GridCell> cells;
  DevBuf<float4> sorted, orig;
  DevBuf<uint32_t> slot_of, rank_of, counters;
  // Two hash tables, occupied-slot lists, counter sets and boxes per map, used alternately: build g fills table g & 1
  // while its scatter launch empties the other one (the slots listed by the build before), so no clear pass is ever on
  // the critical path.  Table t = cells.p + t * table_cap, list t = occ.p + t * occ_cap, set t = counters.p + 8 t,
  // box t = bbox.p + 8 t.
  DevBuf<uint32_t> occ;
  uint32_t table_cap = 0;    // slots allocated per table
  size_t occ_cap = 0;        // entries allocated per list
  int gen = 0;               // build generation
  int cur = 0;               // table of the last build (what view() shows)
  size_t clean_size[2] = {0, 0};  // table t is EMPTY in [0, clean_size[t]) ...
  int filled_n[2] = {0, 0};       // ... unless filled_n[t] != 0: it still holds a build of (at most) that many voxels
  DevBuf<int> bbox;
  DevBuf<float> raw;  // staging of caller bytes for the host-pointer entry points
  // Each map builds on its own stream so that the corner and surf structures of a frame are built concurrently
  // (and their H2D copies overlap); users on the context stream wait on `ready`.
  cudaStream_t stream = nullptr;
  // builds run in line on the context's stream, with no fork / join events: for callers whose builds have nothing to
  // overlap with (the cube map's two search structures sit on the mapping stage's critical path, and every event call is
  // host time there)
  bool in_line = false;
  cudaEvent_t ready = nullptr, ctx_done = nullptr;
  bool pending = false;

  // Add_Points scratch (ikdmap.cu)
  DevBuf<u64> vox_table;
  DevBuf<float4> ins_new, ins_out;
  DevBuf<uint32_t> ins_slot_new, ins_slot_old, ins_keep, ins_pos, ins_bsum;
  int insert_dev(const float* d_src, int n_new, int stride_bytes, int policy, float ds);

  int init(Ctx* c);
  int wait_ready(cudaStream_t user);
  int build_dev(const float* d_src, int n_pts, int stride_bytes, float cell_size);
  int prepare_build(const float* d_src, int n_pts, int stride_bytes, float cell_size, cudaStream_t s, void* job_out);
  // allocate for builds of up to n_pts points now (a buffer that grows later costs a device-wide synchronisation); only
  // while no build is in flight
  int reserve_points(int n_pts);
  const GridCell* cur_cells() const { return cells.p + (size_t)cur * table_cap; }
  const uint32_t* cur_occ() const { return occ.p + (size_t)cur * occ_cap; }
  const uint32_t* cur_counters() const { return counters.p + 8 * cur; }
  int knn_dev(const float* d_q, int nq, int stride_bytes, int k, float max_dist, int32_t* d_idx, float* d_d2);
  GridView view() const;
  void release();
};

// Device-side factor storage (SoA), one slot per stack point (corner slots first).
struct FactorBufs {
  DevBuf<int> type;
  DevBuf<float4> p;    // curr_point (sensor frame), w unused
  DevBuf<double4> a;   // edge: point_a        ; plane: unit normal, w = negative_OA_dot_norm
  DevBuf<double4> b;   // edge: point_b
  DevBuf<int32_t> knn_idx;  // 5 per slot (debug / parity output)
  DevBuf<float> knn_d2;
  int n = 0, nc = 0;
};

struct VgState {      // device-resident scalars of one multi-block VoxelGrid
  int mn[3], mx[3];   // ordered-int encoded float min / max of the finite points
  int n_out;
  int pad;
};

// Front-end scratch and outputs (device-resident between the front end and the registration).
struct FeBufs {
  DevBuf<unsigned char> scanid, picked;
  DevBuf<float> ori, curv;
  DevBuf<int> stats, src_index, label, sort_ind, ring_sharp, ring_lsharp, ring_flat, sharp, lsharp, flat, counts;
  DevBuf<u64> vg_keys;            // multi-block VoxelGrid: sort keys, state, per-chunk head counts / bases
  DevBuf<VgState> vg_state;
  DevBuf<int> vg_chunk;
  DevBuf<uint32_t> vg_hash;         // hash-based block VoxelGrid: global scratch of the (at most two) blocks of a launch
  DevBuf<int> chunk_hist, chunk_base;  // ring counts per 256-point chunk of the frame and their per-ring prefix
  DevBuf<float4> cloud, ring_pts, ring_out, lflat, vox_packed;
  DevBuf<float> raw;                 // staged caller frame
  DevBuf<unsigned char> pc2;         // staged sensor_msgs/PointCloud2 blob (ilsm_pc2_unpack, ilsm_slam_frame_pc2)
  DevBuf<unsigned char> img;         // projection outputs (range | intensity)
  DevBuf<float4> track;
  DevBuf<float4> vox_out;
  DevBuf<int> vox_n;
  int n_in = 0, ring_cap = 0;
};

struct Ctx {
  int device = 0;
  int sm_count = 148;
  cudaStream_t stream = nullptr;
  cudaStream_t aux = nullptr;                       // side stream: work of a frame that does not depend on the odometry
  cudaEvent_t ev_fork = nullptr, ev_join = nullptr;
  std::mutex mu;
  DevBuf<LmState> lm;
  DevBuf<double> partials;  // [blocks][32]
  DevBuf<float> stack_raw;  // staged caller stacks
  DevBuf<int32_t> out_idx;
  DevBuf<float> out_d2;
  PinnedBuf<unsigned char> pinned;
  FactorBufs fac;
  FeBufs fe;
  const int* d_stack_counts = nullptr;  // when set, associate/solve read {nc, ns} from the device (cube-map path)
  int solve_wide = -1;                  // LM solve as a 16-CTA cluster (1) or the portable 8 (0); -1: not probed yet
  DevBuf<float4> rf_lsharp, rf_stack_c, rf_stack_s;  // ilsm_register_frame: less-sharp cloud and the two down-sampled stacks
  DevBuf<int> rf_stack_n;
  Map* qbin = nullptr;        // query binning of the throughput k-NN path (knn_binned.cu): the query cloud grouped by voxel
  DevBuf<int> qwork;          // its work items (4 ints each) + the item counter
  int knn_group = 8;          // queries per group of the binned search (ILSM_KNN_GROUP: 1, 2, 4, 8, 16 or 32)
  int knn_binned_min = 8192;  // query sets at least this large take the binned path (ILSM_KNN_BINNED_MIN overrides)
  size_t partial_blocks = 0;
  bool bulk_attr_set = false;  // normal_eq_bulk_kernel's dynamic shared-memory opt-in done on this device

  int init(int dev);
  void release();
  bool async_build = false;  // ilsm_set_async: host-pointer map builds return without synchronising
  int associate_dev(Map* mc, Map* ms, const float* d_corner, int nc, const float* d_surf, int ns, int stride_bytes,
                    const ilsm_reg_opts& o, bool want_knn, const PoseSrc* src = nullptr);
  int solve_launch(int max_iter, double huber_a, int pass, const PoseDst* dst = nullptr);  // the whole LM solve, one cluster launch
  int odom_associate_dev(Map* mc, Map* ms, const float* d_sharp, int nsh, const float* d_flat, int nfl, int stride_bytes);
  int odometry_dev(Map* mc, Map* ms, const float* d_sharp, int nsh, const float* d_flat, int nfl, int stride_bytes,
                   const ilsm_reg_opts& o);
  // front end (frontend.cu)
  int project_dev(const float* d_cloud, int n, int stride_bytes, unsigned char* d_range, unsigned char* d_inten,
                  float* d_track);
  int features_dev(const float* d_in, int n, int stride_bytes, float min_range);
  int features_launch(const float* d_in, int n, int stride_bytes, float min_range, int ring_cap);  // the 10 plain launches
  cudaGraphExec_t fe_graph_exec = nullptr;  // the captured front-end chain and the parameter hash it was captured for
  unsigned long long fe_graph_key = 0;
  bool fe_graphs_ok = true;                 // false after a failed capture: plain launches from then on
  int pc2_unpack_dev(const unsigned char* d_data, int n, const ilsm_pc2_layout& l, float4* d_out);
  int pc2_pack_dev(const float4* d_in, int n, const ilsm_pc2_layout& l, unsigned char* d_out);
  int voxelgrid_dev(const float* d_in, int n, const int* d_n, int n_slot, int stride_bytes, int ioff, float leaf,
                    float4* d_out, int* d_n_out);
  int voxelgrid_large_dev(const float* d_in, int n, int stride_bytes, int ioff, float leaf, float4* d_out, int* d_n_out,
                          cudaStream_t s = nullptr, int* d_err = nullptr);
  int voxelgrid_pair_dev(const float* d_c, int nc, float leaf_c, float4* d_out_c, const float* d_s, int ns, float leaf_s,
                         float4* d_out_s, int stride_bytes, int ioff, int* d_n_out2, cudaStream_t s, int* d_err = nullptr);
  int gather_dev(const float4* d_cloud, const int* d_idx, const int* d_counts, int slot, int max_n, float4* d_out);
  int register_dev(Map* mc, Map* ms, const float* d_corner, int nc, const float* d_surf, int ns, int stride_bytes,
                   const ilsm_reg_opts& o, const PoseSrc* src = nullptr, const PoseDst* dst = nullptr);
};

// ScanContext keyframe database (one shard when the database is split across ranks).
struct ScQuery;
struct ScDb {
  Ctx* ctx = nullptr;
  int count = 0;
  DevBuf<float> db;        // [count][20][60] float32
  DevBuf<int> bins;        // makeScancontext scratch
  DevBuf<ScQuery> query;
  DevBuf<double> out_dist;
  DevBuf<int> out_id, out_shift;
  DevBuf<u64> part_d;      // per-block top-k lists of the scoring kernel
  DevBuf<int> part_id, part_sh;
  DevBuf<float> stage;     // staged caller descriptors / points
  DevBuf<float> ringkey;   // [count][20] float ring keys (polarcontext_invkeys_mat_, Scancontext.cpp:243,249)
  DevBuf<u64> rk_part;     // ring-key candidate selection: per-block lists + the selected list
  DevBuf<unsigned char> cand_out;  // candidate query outputs (ids | key d2 | distances | shifts)
  // sharded queries (ilsm_sc_init_nccl): per-rank packed top-k records, the all-gathered buffer, the merged result
  void* nccl_comm = nullptr;
  bool nccl_owned = false;
  int nccl_ranks = 1, nccl_rank = 0;
  DevBuf<unsigned char> pk_local, pk_all, pk_out;
  DevBuf<float> qbatch;
  int append_dev(const float* desc, int n_add, bool from_host);
  int query_batch_dev(const float* d_qdesc, int B, int n_search, int id_offset, int k, unsigned char* d_packed);
  // tensor-core prefilter + exact rescoring (scancontext_tc.cu); shards smaller than tc_min keep the plain exact scan
  int tc_min = 4096;
  DevBuf<unsigned char> pf_query;   // PfQuery records of the batch in flight
  DevBuf<float> pf_dist, pf_thr;    // approximate distances [8][n], thresholds [8]
  DevBuf<u64> pf_part, pf_list;     // selection scratch, rescoring lists [8][n]
  DevBuf<int> pf_list_n;
  int query_batch_tc_dev(const float* d_qdesc, int B, int n_search, int id_offset, int k, unsigned char* d_packed, unsigned char* d_shift_dbg);
  int candidates_dev(const float* d_qdesc, int n_search, int num_cand, int* d_id, float* d_key_d2, double* d_dist, int* d_shift);
  int make_dev(const float* d_pts, int n, int stride_bytes, float* d_desc);
  int query_dev(const float* d_qdesc, int n_search, int id_offset, int k, double* d_dist, int* d_id, int* d_shift);
};

}  // namespace ilsm

// the opaque handles of include/ilsm.h
struct ilsm_ctx {
  ilsm::Ctx c;
};
struct ilsm_map {
  ilsm::Map m;
};
struct ilsm_sc {
  ilsm::ScDb d;
};
