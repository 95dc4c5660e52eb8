// ilsm_internal.cuh -- shared device-side building blocks of libilsm_cuda (sm_100a only).
//
// Data layout in HBM:
//   GridCell table  : open-addressing hash of occupied voxels, 16 B/slot {key, start, count}
//   sorted points   : float4 {x, y, z, bits(original index)} grouped by voxel (one 16-B load per candidate)
//   orig points     : float4 {x, y, z, 0} in caller order (neighbour gather for the line/plane fit)
//
// Parity-critical float arithmetic is written with explicit round-to-nearest intrinsics so that no FMA
// contraction can happen whatever the compiler flags: the reference runs on x86-64 SSE2 without FMA
// (CMakeLists.txt:5-6) and k-NN indices must match it bit for bit.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace ilsm {

typedef unsigned long long u64;

constexpr u64 kEmptyKey = 0xFFFFFFFFFFFFFFFFull;
constexpr u64 kSentinel = 0xFFFFFFFFFFFFFFFFull;  // "no candidate" k-NN key (d2 bits = NaN pattern, idx = -1)
constexpr int kCoordOff = 1 << 20;                // voxel coordinates are biased into [0, 2^21)
constexpr int kCoordLim = (1 << 20) - 4;
constexpr int kMaxRing = 6;                       // rings searched before the brute-force fallback

struct __align__(16) GridCell {
  u64 key;
  uint32_t start;
  uint32_t count;
};

struct GridView {
  const GridCell* cells;
  const float4* sorted;
  const float4* orig;
  const int* bbox;  // device int[6]: min cx,cy,cz / max cx,cy,cz of occupied voxels
  uint32_t mask;
  int log2_size;
  float cell, inv_cell;
  int n;
};

// Programmatic dependent launch (sm_90+): a kernel launched with launch_pdl() may be scheduled while its predecessor
// in the stream is still running; it must not touch anything the predecessor produces before pdl_wait(), which
// returns once the predecessor has completed and its writes are visible.  pdl_launch_dependents() lets the NEXT
// kernel's launch overlap this one.  Both are no-ops for ordinary launches.  Every kernel launched through
// launch_pdl() calls pdl_entry() first.
__device__ __forceinline__ void pdl_entry() {
  asm volatile("griddepcontrol.wait;" ::: "memory");
  asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
}

__device__ __forceinline__ int voxel_coord(float x, float inv_cell) { return __float2int_rd(__fmul_rn(x, inv_cell)); }

__device__ __forceinline__ u64 pack_voxel(int cx, int cy, int cz) {
  return ((u64)(uint32_t)(cx + kCoordOff) << 42) | ((u64)(uint32_t)(cy + kCoordOff) << 21) |
         (u64)(uint32_t)(cz + kCoordOff);
}

__device__ __forceinline__ uint32_t hash_voxel(u64 key, int log2_size) {
  return (uint32_t)((key * 0x9E3779B97F4A7C15ull) >> (64 - log2_size));
}

// FLANN L2_Simple<float> / ikd-Tree calc_dist: ((dx*dx)+(dy*dy))+(dz*dz), round-to-nearest, no FMA.
__device__ __forceinline__ float dist2_rn(float qx, float qy, float qz, float px, float py, float pz) {
  float dx = __fsub_rn(qx, px), dy = __fsub_rn(qy, py), dz = __fsub_rn(qz, pz);
  return __fadd_rn(__fadd_rn(__fmul_rn(dx, dx), __fmul_rn(dy, dy)), __fmul_rn(dz, dz));
}

// (d2, idx) packed so that unsigned 64-bit order == (ascending d2, ascending idx): d2 >= +0 so its IEEE bits
// are monotone as an unsigned integer.
__device__ __forceinline__ u64 pack_cand(float d2, uint32_t idx) { return ((u64)__float_as_uint(d2) << 32) | idx; }
__device__ __forceinline__ float cand_d2(u64 k) { return __uint_as_float((uint32_t)(k >> 32)); }
__device__ __forceinline__ int cand_idx(u64 k) { return (int)(uint32_t)(k & 0xFFFFFFFFull); }

__device__ __forceinline__ uint4 ldg_cell(const GridCell* c) { return __ldg(reinterpret_cast<const uint4*>(c)); }

// Look a voxel up; returns count (0 when absent) and sets start.
__device__ __forceinline__ uint32_t probe_voxel(const GridView& g, u64 key, uint32_t& start) {
  uint32_t slot = hash_voxel(key, g.log2_size);
#pragma unroll 1
  for (;;) {
    uint4 e = ldg_cell(g.cells + slot);
    u64 k = ((u64)e.y << 32) | e.x;
    if (k == key) {
      start = e.z;
      return e.w;
    }
    if (k == kEmptyKey) return 0;
    slot = (slot + 1) & g.mask;
  }
}

// Per-lane candidate list: K best (d2, idx) keys in ascending order, optionally with the candidates' coordinates
// (the association kernel needs the 5 neighbours' xyz for the fit and saves a dependent gather by carrying them).
template <int K, bool XYZ>
struct CandList {
  u64 key[K];
  float x[XYZ ? K : 1], y[XYZ ? K : 1], z[XYZ ? K : 1];
  __device__ __forceinline__ void clear() {
#pragma unroll
    for (int k = 0; k < K; ++k) key[k] = kSentinel;
  }
  __device__ __forceinline__ void insert(u64 kk, float px, float py, float pz) {
    if (kk < key[K - 1]) {
      key[K - 1] = kk;
      if (XYZ) x[K - 1] = px, y[K - 1] = py, z[K - 1] = pz;
#pragma unroll
      for (int s = K - 1; s > 0; --s) {
        const bool sw = key[s] < key[s - 1];
        const u64 a = key[s - 1], b = key[s];
        key[s - 1] = sw ? b : a;
        key[s] = sw ? a : b;
        if (XYZ) {
          const float ax = x[s - 1], bx = x[s], ay = y[s - 1], by = y[s], az = z[s - 1], bz = z[s];
          x[s - 1] = sw ? bx : ax, x[s] = sw ? ax : bx;
          y[s - 1] = sw ? by : ay, y[s] = sw ? ay : by;
          z[s - 1] = sw ? bz : az, z[s] = sw ? az : bz;
        }
      }
    }
  }
  __device__ __forceinline__ void pop_front() {
#pragma unroll
    for (int s = 0; s < K - 1; ++s) {
      key[s] = key[s + 1];
      if (XYZ) x[s] = x[s + 1], y[s] = y[s + 1], z[s] = z[s + 1];
    }
    key[K - 1] = kSentinel;
  }
};

// Group-uniform result of a search: sorted keys (+ coordinates when requested).
template <int K, bool XYZ>
struct KnnResult {
  u64 key[K];
  float x[XYZ ? K : 1], y[XYZ ? K : 1], z[XYZ ? K : 1];
};

// Per-warp shared-memory scratch of the search: the candidate ranges found by the 32 lanes of one probing round.
struct WarpScratch {
  uint32_t start[32];
  uint32_t prefix[33];
};

__device__ __forceinline__ uint32_t redux_min_u32(unsigned mask, uint32_t v) {
  uint32_t r;
  asm volatile("redux.sync.min.u32 %0, %1, %2;" : "=r"(r) : "r"(v), "r"(mask));
  return r;
}

// Merge the per-lane sorted lists of a warp into the warp-uniform sorted result: K rounds of a 64-bit arg-min
// done as two 32-bit REDUX ops (high word = d2 bits, low word = index).
template <int K, bool XYZ>
__device__ __forceinline__ void warp_merge(CandList<K, XYZ>& best, KnnResult<K, XYZ>& res) {
#pragma unroll
  for (int k = 0; k < K; ++k) {
    const uint32_t hi = (uint32_t)(best.key[0] >> 32), lo = (uint32_t)best.key[0];
    const uint32_t mhi = redux_min_u32(0xffffffffu, hi);
    const uint32_t mlo = redux_min_u32(0xffffffffu, hi == mhi ? lo : 0xFFFFFFFFu);
    const u64 m = ((u64)mhi << 32) | mlo;
    res.key[k] = m;
    const bool mine = best.key[0] == m && m != kSentinel;  // keys are unique: at most one lane
    if (XYZ) {
      const unsigned who = __ballot_sync(0xffffffffu, mine);
      const int src = who ? __ffs(who) - 1 : 0;
      res.x[k] = __shfl_sync(0xffffffffu, best.x[0], src);
      res.y[k] = __shfl_sync(0xffffffffu, best.y[0], src);
      res.z[k] = __shfl_sync(0xffffffffu, best.z[0], src);
    }
    if (mine) best.pop_front();
  }
}

// Exact K-NN of (qx,qy,qz) by one warp.
// Ring expansion over the voxel hash.  Each probing round gives every lane one voxel; the (start,count) pairs go
// through a warp prefix sum into shared memory and the CANDIDATE POINTS (not the voxels) are then dealt out to the
// lanes, so that all point loads of a round are in flight together and the work is balanced whatever the voxel
// occupancy.  The first pass visits the 3x3x3 block around the query voxel, further passes add one Chebyshev
// shell each.  The search stops when (a) the K-th distance is provably smaller than the distance to anything
// unvisited, (b) the visited block already contains the ball of radius sqrt(max_d2) (results beyond it are
// "don't care"), or (c) the block covers every occupied voxel.  After kMaxRing rings it falls back to a
// coalesced brute-force sweep of the whole map (same candidate loop, one range).  res is identical in all lanes.
// bb = bounding box of occupied voxels (preloaded by the caller so that its latency overlaps the query's).
template <int K, bool XYZ>
__device__ __forceinline__ void knn_search(const GridView& g, const int (&bb)[6], float qx, float qy, float qz,
                                           float max_d2, unsigned lane, WarpScratch& ws, KnnResult<K, XYZ>& res) {
  CandList<K, XYZ> best;
  best.clear();
#pragma unroll
  for (int k = 0; k < K; ++k) res.key[k] = kSentinel;
  if (g.n <= 0) return;

  const float ux = __fmul_rn(qx, g.inv_cell), uy = __fmul_rn(qy, g.inv_cell), uz = __fmul_rn(qz, g.inv_cell);
  // query outside the addressable voxel range (or NaN): brute force keeps the result exact
  const bool in_range = fabsf(ux) < (float)kCoordLim && fabsf(uy) < (float)kCoordLim && fabsf(uz) < (float)kCoordLim;
  const int cx = __float2int_rd(ux), cy = __float2int_rd(uy), cz = __float2int_rd(uz);
  const float fx = ux - (float)cx, fy = uy - (float)cy, fz = uz - (float)cz;
  const float fmin = fminf(fminf(fminf(fx, 1.f - fx), fminf(fy, 1.f - fy)), fminf(fz, 1.f - fz));
  const float umax = fmaxf(fmaxf(fabsf(ux), fabsf(uy)), fabsf(uz));

  int r = in_range ? 1 : kMaxRing + 1;
#pragma unroll 1
  for (;;) {
    const bool brute = r > kMaxRing;
    const int side = 2 * r + 1, total = brute ? 1 : side * side * side;
#pragma unroll 1
    for (int base = 0; base < total; base += 32) {
      const int t = base + (int)lane;
      uint32_t start = 0, cnt = 0;
      if (brute) {
        if (lane == 0) cnt = (uint32_t)g.n;
      } else if (t < total) {
        int dz = t / (side * side), rem = t - dz * side * side;
        int dy = rem / side, dx = rem - dy * side;
        dx -= r, dy -= r, dz -= r;
        const bool interior = r > 1 && abs(dx) < r && abs(dy) < r && abs(dz) < r;  // visited by earlier passes
        const int vx = cx + dx, vy = cy + dy, vz = cz + dz;
        if (!interior && vx >= bb[0] && vx <= bb[3] && vy >= bb[1] && vy <= bb[4] && vz >= bb[2] && vz <= bb[5])
          cnt = probe_voxel(g, pack_voxel(vx, vy, vz), start);
      }
      uint32_t inc = cnt;
#pragma unroll
      for (int off = 1; off < 32; off <<= 1) {
        const uint32_t v = __shfl_up_sync(0xffffffffu, inc, off);
        if (lane >= (unsigned)off) inc += v;
      }
      const uint32_t ncand = __shfl_sync(0xffffffffu, inc, 31);
      ws.start[lane] = start;
      ws.prefix[lane + 1] = inc;
      if (lane == 0) ws.prefix[0] = 0;
      __syncwarp();
#pragma unroll 1
      for (uint32_t k = lane; k < ncand; k += 32) {
        int c = 0;  // largest c with prefix[c] <= k: the non-empty range that contains candidate k
#pragma unroll
        for (int step = 16; step > 0; step >>= 1)
          if (ws.prefix[c + step] <= k) c += step;
        const float4 p = __ldg(g.sorted + ws.start[c] + (k - ws.prefix[c]));
        best.insert(pack_cand(dist2_rn(qx, qy, qz, p.x, p.y, p.z), __float_as_uint(p.w)), p.x, p.y, p.z);
      }
      __syncwarp();
    }
    warp_merge<K, XYZ>(best, res);
    if (brute) return;

    // distance (metres) from the query to the nearest unvisited voxel, made conservative against the float
    // rounding of the voxel coordinates of both the query and any map point (2^-24 relative each).
    const float margin = (umax + (float)r + 2.f) * 2.4e-7f;
    const float bound = ((float)r + fmin - margin) * g.cell;
    const float b2 = bound > 0.f ? bound * bound * 0.999999f : 0.f;
    const float dk = cand_d2(res.key[K - 1]);  // NaN pattern when fewer than K found
    const bool done = (res.key[K - 1] != kSentinel && dk < b2) || (max_d2 > 0.f && b2 >= max_d2) ||
                      (cx - r <= bb[0] && cx + r >= bb[3] && cy - r <= bb[1] && cy + r >= bb[4] && cz - r <= bb[2] &&
                       cz + r >= bb[5]);
    if (done) return;
    ++r;
    // re-seed: lane 0 carries the merged result forward (brute force restarts from scratch), others start empty
    best.clear();
    if (lane == 0 && r <= kMaxRing) {
#pragma unroll
      for (int k = 0; k < K; ++k) {
        best.key[k] = res.key[k];
        if (XYZ) best.x[k] = res.x[k], best.y[k] = res.y[k], best.z[k] = res.z[k];
      }
    }
  }
}

__device__ __forceinline__ void load_bbox(const GridView& g, int (&bb)[6]) {
#pragma unroll
  for (int i = 0; i < 6; ++i) bb[i] = __ldg(g.bbox + i);
}

// ------------------------------------------------------------------------------------------------
// fast fp64 reciprocal / rsqrt: MUFU seed (2^-23 relative) + two Newton steps => ~1 ulp, ~3x shorter dependent
// chain than the IEEE sequences; used only where the result is compared against the oracle with a tolerance
// (fits, residuals, LM), never in the float k-NN distance path.
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ double frcp(double x) {
  double r;
  asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(r) : "d"(x));
  r = fma(r, fma(-x, r, 1.0), r);
  r = fma(r, fma(-x, r, 1.0), r);
  return r;
}
__device__ __forceinline__ double frsqrt(double x) {
  double y;
  asm("rsqrt.approx.ftz.f64 %0, %1;" : "=d"(y) : "d"(x));
  const double h = 0.5 * x;
  y = y * fma(-(h * y), y, 1.5);
  y = y * fma(-(h * y), y, 1.5);
  return y;
}
__device__ __forceinline__ double fsqrt(double x) { return x > 0.0 ? x * frsqrt(x) : 0.0; }

// ------------------------------------------------------------------------------------------------
// double-precision helpers (Eigen-compatible operation order, no FMA)
// ------------------------------------------------------------------------------------------------
struct D3 {
  double x, y, z;
};
__device__ __forceinline__ D3 d3(double x, double y, double z) { return D3{x, y, z}; }
__device__ __forceinline__ double dmul(double a, double b) { return __dmul_rn(a, b); }
__device__ __forceinline__ double dadd(double a, double b) { return __dadd_rn(a, b); }
__device__ __forceinline__ double dsub(double a, double b) { return __dsub_rn(a, b); }
__device__ __forceinline__ D3 cross_rn(D3 a, D3 b) {
  return D3{dsub(dmul(a.y, b.z), dmul(a.z, b.y)), dsub(dmul(a.z, b.x), dmul(a.x, b.z)),
            dsub(dmul(a.x, b.y), dmul(a.y, b.x))};
}
// Eigen QuaternionBase::_transformVector: uv = 2 (u x v);  v + w*uv + u x uv    (q = x,y,z,w)
__device__ __forceinline__ D3 quat_rotate(const double q[4], D3 v) {
  D3 u{q[0], q[1], q[2]};
  D3 uv = cross_rn(u, v);
  uv = D3{dadd(uv.x, uv.x), dadd(uv.y, uv.y), dadd(uv.z, uv.z)};
  D3 c = cross_rn(u, uv);
  return D3{dadd(dadd(v.x, dmul(q[3], uv.x)), c.x), dadd(dadd(v.y, dmul(q[3], uv.y)), c.y),
            dadd(dadd(v.z, dmul(q[3], uv.z)), c.z)};
}

}  // namespace ilsm
