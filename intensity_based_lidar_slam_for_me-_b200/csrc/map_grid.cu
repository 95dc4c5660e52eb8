// map_grid.cu -- K1: voxel-hashed local map (build) and the stand-alone exact k-NN kernel.
//
// Replaces pcl::KdTreeFLANN::setInputCloud/nearestKSearch (laserMapping.cpp:631-634,673,753;
// laserOdometry.cpp:452,574,807-808) and ikd-Tree Build/Nearest_Search (mapOptimization.cpp:192,393).
//
// Build = 4 short kernels, no sort:
//   clear   : the slots the PREVIOUS build occupied <- EMPTY (its occupied-slot list), counters <- 0; a full clear of
//             the table only when the table grew or was reallocated
//   count   : every point claims its voxel slot with atomicCAS and takes a rank with atomicAdd; slots claimed for the
//             first time are appended to the occupied-slot list (one atomicAdd per block)
//   alloc   : every occupied slot (the list, not the table) gets a contiguous range (one atomicAdd per block)
//   scatter : points are written as float4 {x,y,z,bits(index)} into their voxel's range
// The table has >= 2 N slots (the only safe bound before the points are seen) but only the occupied voxels -- a few
// per cent of it for a dense map -- are ever touched again: at N = 2 M the full clear + full-table alloc pass moved
// 190 MB of the build's 480 MB.
// The memory order of voxels/points is not deterministic, the k-NN result is: selection uses the total
// order (d2, original index).
#include "ilsm_host.hpp"

namespace ilsm {

__global__ void grid_clear_kernel(GridCell* cells, uint32_t size, int* bbox, uint32_t* counters) {
  pdl_entry();
  uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < size) {
    uint4 e;
    e.x = 0xFFFFFFFFu, e.y = 0xFFFFFFFFu, e.z = 0u, e.w = 0u;
    reinterpret_cast<uint4*>(cells)[i] = e;
  }
  if (i < 3) bbox[i] = INT_MAX;
  if (i >= 3 && i < 6) bbox[i] = INT_MIN;
  if (i < 6) counters[i] = 0;  // [0] cursor, [2] skipped points, [4], [5] occupied voxels (ping-pong by build generation)
}

// counters: [0] cursor, [2] skipped points, [4 + (gen & 1)] occupied voxels of build `gen`
__global__ void grid_clear_sparse_kernel(GridCell* cells, const uint32_t* __restrict__ occ, int* bbox, uint32_t* counters,
                                         int gen) {
  pdl_entry();
  const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  const uint32_t n_prev = counters[4 + ((gen - 1) & 1)];  // nobody writes this one in this launch
  if (i < n_prev) {
    uint4 e;
    e.x = 0xFFFFFFFFu, e.y = 0xFFFFFFFFu, e.z = 0u, e.w = 0u;
    reinterpret_cast<uint4*>(cells)[occ[i]] = e;
  }
  if (i < 3) bbox[i] = INT_MAX;
  if (i >= 3 && i < 6) bbox[i] = INT_MIN;
  if (i < 3) counters[i] = 0;
  if (i == 3) counters[4 + (gen & 1)] = 0;
}

__device__ __forceinline__ bool load_point(const float* src, int stride_f, int i, float& x, float& y, float& z) {
  const float* p = src + (size_t)i * stride_f;
  x = __ldg(p), y = __ldg(p + 1), z = __ldg(p + 2);
  return isfinite(x) && isfinite(y) && isfinite(z);
}

__global__ void grid_count_kernel(const float* __restrict__ src, int n, int stride_f, int ioff, float inv_cell, GridCell* cells,
                                  uint32_t mask, int log2_size, float4* __restrict__ orig,
                                  uint32_t* __restrict__ slot_of, uint32_t* __restrict__ rank_of,
                                  uint32_t* counters, uint32_t* __restrict__ occ, int gen) {
  pdl_entry();
  __shared__ uint32_t s_wnew[8];
  __shared__ uint32_t s_obase;
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  const int lane = threadIdx.x & 31;
  const bool in = i < n;
  float x = 0.f, y = 0.f, z = 0.f;
  bool ok = in && load_point(src, stride_f, i, x, y, z);
  // w keeps the caller's intensity channel (LOAM clouds carry scanID + 0.1*relTime there, laserOdometry.cpp:461)
  if (in) orig[i] = make_float4(x, y, z, ioff >= 0 ? __ldg(src + (size_t)i * stride_f + ioff) : 0.f);
  int cx = 0, cy = 0, cz = 0;
  if (ok) {
    float ux = __fmul_rn(x, inv_cell), uy = __fmul_rn(y, inv_cell), uz = __fmul_rn(z, inv_cell);
    ok = fabsf(ux) < (float)kCoordLim && fabsf(uy) < (float)kCoordLim && fabsf(uz) < (float)kCoordLim;
    cx = __float2int_rd(ux), cy = __float2int_rd(uy), cz = __float2int_rd(uz);
  }
  if (in && !ok) {
    slot_of[i] = 0xFFFFFFFFu;
    atomicAdd(&counters[2], 1u);
  }
  // warp-aggregated claim: map clouds arrive voxel-ordered (VoxelGrid output, cube by cube), so the lanes of a warp
  // mostly share a handful of voxels -- one atomicCAS + one atomicAdd per distinct voxel of the warp instead of one
  // per point.  All 32 lanes take part in the match; lanes without a point carry a key nobody shares.
  const u64 key = ok ? pack_voxel(cx, cy, cz) : (kEmptyKey - 1ull - (u64)lane);
  const unsigned peers = __match_any_sync(0xffffffffu, key);
  bool created = false;  // this lane claimed a slot nobody had claimed before
  uint32_t my_slot = 0;
  if (ok) {
    const int leader = __ffs(peers) - 1;
    uint32_t slot = 0, base = 0;
    if (lane == leader) {
      slot = hash_voxel(key, log2_size);
      for (;;) {
        u64 prev = atomicCAS(&cells[slot].key, kEmptyKey, key);
        if (prev == kEmptyKey) created = true;
        if (prev == kEmptyKey || prev == key) break;
        slot = (slot + 1) & mask;
      }
      base = atomicAdd(&cells[slot].count, (uint32_t)__popc(peers));
      my_slot = slot;
    }
    slot = __shfl_sync(peers, slot, leader);
    base = __shfl_sync(peers, base, leader);
    slot_of[i] = slot;
    rank_of[i] = base + (uint32_t)__popc(peers & ((1u << lane) - 1u));
  }
  // occupied-slot list: block-wide count of newly claimed slots, ONE atomicAdd on the list cursor per block
  const unsigned newb = __ballot_sync(0xffffffffu, created);
  const int warp = threadIdx.x >> 5;
  if (lane == 0) s_wnew[warp] = (uint32_t)__popc(newb);
  __syncthreads();
  if (threadIdx.x == 0) {
    uint32_t tot = 0;
#pragma unroll
    for (int w = 0; w < 8; ++w) tot += s_wnew[w];
    s_obase = tot ? atomicAdd(&counters[4 + (gen & 1)], tot) : 0u;
  }
  __syncthreads();
  if (created) {
    uint32_t wb = 0;
#pragma unroll
    for (int w = 0; w < 8; ++w)
      if (w < warp) wb += s_wnew[w];
    occ[s_obase + wb + (uint32_t)__popc(newb & ((1u << lane) - 1u))] = my_slot;
  }
}

// Every occupied slot gets a contiguous range of the sorted array (warp-aggregated atomicAdd on one cursor) and
// the bounding box of occupied voxels is reduced per block (6 atomics per block instead of 6 per voxel).
__global__ void grid_alloc_kernel(GridCell* cells, const uint32_t* __restrict__ occ, uint32_t* counters, int* bbox, int gen) {
  pdl_entry();
  const uint32_t li = blockIdx.x * blockDim.x + threadIdx.x;
  const uint32_t size = counters[4 + (gen & 1)];  // occupied voxels of this build; the grid covers the upper bound n
  uint32_t cnt = 0;
  int lo[3] = {INT_MAX, INT_MAX, INT_MAX}, hi[3] = {INT_MIN, INT_MIN, INT_MIN};
  const uint32_t i = li < size ? occ[li] : 0u;
  if (li < size) {
    uint4 e = *reinterpret_cast<const uint4*>(cells + i);
    cnt = e.w;
    if (cnt) {
      u64 key = ((u64)e.y << 32) | e.x;
      int c[3] = {(int)((key >> 42) & 0x1FFFFF) - kCoordOff, (int)((key >> 21) & 0x1FFFFF) - kCoordOff,
                  (int)(key & 0x1FFFFF) - kCoordOff};
#pragma unroll
      for (int a = 0; a < 3; ++a) lo[a] = hi[a] = c[a];
    }
  }
  // block-wide exclusive prefix of the counts, ONE atomicAdd on the cursor per block
  __shared__ uint32_t s_wsum[8];
  __shared__ uint32_t s_base;
  const uint32_t lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  uint32_t inc = cnt;
#pragma unroll
  for (int off = 1; off < 32; off <<= 1) {
    uint32_t v = __shfl_up_sync(0xffffffffu, inc, off);
    if (lane >= (uint32_t)off) inc += v;
  }
  const uint32_t total = __shfl_sync(0xffffffffu, inc, 31);
  if (lane == 31) s_wsum[warp] = inc;
  __syncthreads();
  uint32_t wbase = 0, btotal = 0;
#pragma unroll
  for (int w = 0; w < 8; ++w) {
    if (w < (int)warp) wbase += s_wsum[w];
    btotal += s_wsum[w];
  }
  if (threadIdx.x == 0) s_base = btotal ? atomicAdd(&counters[0], btotal) : 0u;
  __syncthreads();
  if (li < size && cnt) cells[i].start = s_base + wbase + inc - cnt;
  // bounding box
#pragma unroll
  for (int a = 0; a < 3; ++a) {
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) {
      lo[a] = min(lo[a], __shfl_xor_sync(0xffffffffu, lo[a], off));
      hi[a] = max(hi[a], __shfl_xor_sync(0xffffffffu, hi[a], off));
    }
  }
  __shared__ int s_lo[3], s_hi[3];
  if (threadIdx.x < 3) s_lo[threadIdx.x] = INT_MAX, s_hi[threadIdx.x] = INT_MIN;
  __syncthreads();
  if (lane == 0 && total) {
#pragma unroll
    for (int a = 0; a < 3; ++a) atomicMin(&s_lo[a], lo[a]), atomicMax(&s_hi[a], hi[a]);
  }
  __syncthreads();
  if (threadIdx.x < 3 && s_lo[threadIdx.x] != INT_MAX) {
    atomicMin(&bbox[threadIdx.x], s_lo[threadIdx.x]);
    atomicMax(&bbox[3 + threadIdx.x], s_hi[threadIdx.x]);
  }
}

__global__ void grid_scatter_kernel(const float4* __restrict__ orig, int n, const GridCell* __restrict__ cells,
                                    const uint32_t* __restrict__ slot_of, const uint32_t* __restrict__ rank_of,
                                    float4* __restrict__ sorted) {
  pdl_entry();
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  uint32_t slot = slot_of[i];
  if (slot == 0xFFFFFFFFu) return;
  float4 p = orig[i];
  p.w = __uint_as_float((uint32_t)i);
  sorted[cells[slot].start + rank_of[i]] = p;
}

// ---------------------------------------------------------------------------------------------------
// stand-alone k-NN kernel: one G-lane group per query
// ---------------------------------------------------------------------------------------------------
template <int K>
__global__ void __launch_bounds__(128) knn_kernel(GridView g, const float* __restrict__ q, int nq, int stride_f, int k_out,
                                                  float max_d2, int32_t* __restrict__ idx, float* __restrict__ d2) {
  pdl_entry();
  __shared__ WarpScratch scratch[4];
  const unsigned lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  int bb[6];
  load_bbox(g, bb);
  const int nwarps = gridDim.x * 4;
  for (int gid = blockIdx.x * 4 + warp; gid < nq; gid += nwarps) {
    const float* qp = q + (size_t)gid * stride_f;
    const float qx = __ldg(qp), qy = __ldg(qp + 1), qz = __ldg(qp + 2);
    KnnResult<K, false> res;
    knn_search<K, false>(g, bb, qx, qy, qz, max_d2, lane, scratch[warp], res);
    if (lane == 0) {
#pragma unroll
      for (int k = 0; k < K; ++k) {
        if (k < k_out) {
          bool have = res.key[k] != kSentinel;
          idx[(size_t)gid * k_out + k] = have ? cand_idx(res.key[k]) : -1;
          d2[(size_t)gid * k_out + k] = have ? cand_d2(res.key[k]) : __int_as_float(0x7f800000);
        }
      }
    }
  }
}

int knn_binned_dev(Ctx* ctx, Map* m, const float* d_q, int nq, int stride_bytes, int k, float max_dist, int32_t* d_idx, float* d_d2);

static inline int ilog2_ceil(uint32_t v) {
  int l = 0;
  while ((1u << l) < v) ++l;
  return l;
}

int Map::init(Ctx* c) {
  ctx = c;
  ILSM_CUDA(cudaStreamCreateWithFlags(&stream, cudaStreamNonBlocking));
  ILSM_CUDA(cudaEventCreateWithFlags(&ready, cudaEventDisableTiming));
  ILSM_CUDA(cudaEventCreateWithFlags(&ctx_done, cudaEventDisableTiming));
  return ILSM_OK;
}

int Map::wait_ready(cudaStream_t user) {
  if (pending) {
    ILSM_CUDA(cudaStreamWaitEvent(user, ready, 0));
    if (user == ctx->stream) pending = false;
  }
  return ILSM_OK;
}

int Map::build_dev(const float* d_src, int n_pts, int stride_bytes, float cell_size) {
  if (n_pts < 0 || (stride_bytes % 4) != 0 || stride_bytes < 12) return fail(ILSM_ERR_INVALID_ARG, "map_build: bad n/stride");
  if (!(cell_size > 0.f)) cell_size = 1.0f;
  cudaStream_t s = stream;
  // order this (re)build after everything already enqueued on the context stream (previous users of the map,
  // producers of d_src), then run it on the map's own stream
  ILSM_CUDA(cudaEventRecord(ctx_done, ctx->stream));
  ILSM_CUDA(cudaStreamWaitEvent(s, ctx_done, 0));
  n = n_pts;
  cell = cell_size;
  inv_cell = 1.0f / cell_size;
  int want_log2 = ilog2_ceil((uint32_t)(n_pts > 0 ? 2u * (uint32_t)n_pts : 2u));
  if (want_log2 < 10) want_log2 = 10;
  uint32_t want_size = 1u << want_log2;
  const GridCell* cells_before = cells.p;
  const uint32_t *occ_before = occ.p, *counters_before = counters.p;  // a reallocated list / counter loses the last build's record
  int rc;
  if ((rc = cells.reserve(want_size)) || (rc = sorted.reserve(n_pts + 1)) || (rc = orig.reserve(n_pts + 1)) ||
      (rc = slot_of.reserve(n_pts + 1)) || (rc = rank_of.reserve(n_pts + 1)) || (rc = bbox.reserve(8)) ||
      (rc = counters.reserve(8)) || (rc = occ.reserve(n_pts + 1)))
    return rc;
  log2_size = want_log2;
  table_size = want_size;
  const int T = 256;
  // Invariant: every slot of cells[0, clean_size) is EMPTY except the ones listed in occ[0, n_occ(previous build)).
  // A build therefore only has to clear that list -- unless the table is new, or grew beyond the clean prefix.
  const bool full_clear = gen == 0 || cells.p != cells_before || occ.p != occ_before || counters.p != counters_before ||
                          (size_t)table_size > clean_size;
  if (full_clear) {
    ILSM_CUDA(launch_pdl(grid_clear_kernel, dim3((table_size + T - 1) / T), dim3(T), 0, s, cells.p, table_size, bbox.p, counters.p));
    clean_size = table_size;
  } else {
    const int cover = prev_n > 8 ? prev_n : 8;  // the previous build's point count bounds its occupied-voxel count
    ILSM_CUDA(launch_pdl(grid_clear_sparse_kernel, dim3((cover + T - 1) / T), dim3(T), 0, s, cells.p, (const uint32_t*)occ.p, bbox.p,
                         counters.p, gen));
  }
  if (n_pts > 0) {
    const int ioff = stride_bytes >= 32 ? 4 : (stride_bytes >= 16 ? 3 : -1);
    ILSM_CUDA(launch_pdl(grid_count_kernel, dim3((n_pts + T - 1) / T), dim3(T), 0, s, d_src, n_pts, stride_bytes / 4, ioff, inv_cell,
                         cells.p, table_size - 1, log2_size, orig.p, slot_of.p, rank_of.p, counters.p, occ.p, gen));
    ILSM_CUDA(launch_pdl(grid_alloc_kernel, dim3((n_pts + T - 1) / T), dim3(T), 0, s, cells.p, (const uint32_t*)occ.p, counters.p, bbox.p,
                         gen));
    ILSM_CUDA(launch_pdl(grid_scatter_kernel, dim3((n_pts + T - 1) / T), dim3(T), 0, s, (const float4*)orig.p, n_pts,
                         (const GridCell*)cells.p, (const uint32_t*)slot_of.p, (const uint32_t*)rank_of.p, sorted.p));
  }
  prev_n = n_pts;
  gen += 1;
  count_launches(n_pts > 0 ? 4 : 1);
  ILSM_CUDA(cudaEventRecord(ready, s));
  pending = true;
  return check_launch("map_build");
}

GridView Map::view() const {
  GridView g;
  g.cells = cells.p;
  g.sorted = sorted.p;
  g.orig = orig.p;
  g.bbox = bbox.p;
  g.mask = table_size - 1;
  g.log2_size = log2_size;
  g.cell = cell;
  g.inv_cell = inv_cell;
  g.n = n;
  return g;
}

template <int K>
static int launch_knn(const GridView& g, const float* d_q, int nq, int stride_f, int k, float max_d2, int32_t* d_idx,
                       float* d_d2, cudaStream_t s, int sm_count) {
  // one warp per query; at most ~16 resident warps per SM, further queries are taken grid-stride
  long long blocks = ((long long)nq + 3) / 4, cap = (long long)sm_count * 4 * 4;
  if (blocks > cap) blocks = cap;
  ILSM_CUDA(launch_pdl(knn_kernel<K>, dim3((unsigned)blocks), dim3(128), 0, s, g, d_q, nq, stride_f, k, max_d2, d_idx, d_d2));
  return ILSM_OK;
}

int Map::knn_dev(const float* d_q, int nq, int stride_bytes, int k, float max_dist, int32_t* d_idx, float* d_d2) {
  if (nq < 0 || k < 1 || k > 8 || (stride_bytes % 4) != 0 || stride_bytes < 12)
    return fail(ILSM_ERR_INVALID_ARG, "knn: bad nq/k/stride");
  if (table_size == 0) return fail(ILSM_ERR_STATE, "knn: map not built");
  if (nq == 0) return ILSM_OK;
  // throughput regime: bin the queries by voxel and serve each group with one warp (knn_binned.cu); small query sets are
  // latency-bound and keep one warp per query
  if (nq >= ctx->knn_binned_min && n > 0) return knn_binned_dev(ctx, this, d_q, nq, stride_bytes, k, max_dist, d_idx, d_d2);
  GridView g = view();
  float max_d2 = max_dist > 0.f ? max_dist * max_dist : 0.f;
  cudaStream_t s = ctx->stream;
  {
    int rc = wait_ready(s);
    if (rc) return rc;
  }
  int stride_f = stride_bytes / 4;
  int rc;
  if (k == 1)
    rc = launch_knn<1>(g, d_q, nq, stride_f, k, max_d2, d_idx, d_d2, s, ctx->sm_count);
  else if (k <= 5)
    rc = launch_knn<5>(g, d_q, nq, stride_f, k, max_d2, d_idx, d_d2, s, ctx->sm_count);
  else
    rc = launch_knn<8>(g, d_q, nq, stride_f, k, max_d2, d_idx, d_d2, s, ctx->sm_count);
  if (rc) return rc;
  count_launches(1);
  return check_launch("knn");
}

}  // namespace ilsm
