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

template <int K>
__device__ __forceinline__ void topk_insert(u64 (&best)[K], u64 key) {
  if (key < best[K - 1]) {
    best[K - 1] = key;
#pragma unroll
    for (int s = K - 1; s > 0; --s) {
      u64 a = best[s - 1], b = best[s];
      bool sw = b < a;
      best[s - 1] = sw ? b : a;
      best[s] = sw ? a : b;
    }
  }
}

template <int K>
__device__ __forceinline__ void scan_voxel(const GridView& g, int cx, int cy, int cz, float qx, float qy, float qz,
                                           u64 (&best)[K]) {
  uint32_t start;
  uint32_t cnt = probe_voxel(g, pack_voxel(cx, cy, cz), start);
  for (uint32_t j = 0; j < cnt; ++j) {
    float4 p = __ldg(g.sorted + start + j);
    float d2 = dist2_rn(qx, qy, qz, p.x, p.y, p.z);
    topk_insert<K>(best, pack_cand(d2, __float_as_uint(p.w)));
  }
}

// Merge the per-lane sorted lists of a G-lane group into the group-uniform sorted result.
template <int K, int G>
__device__ __forceinline__ void group_merge(u64 (&best)[K], u64 (&res)[K], unsigned gmask) {
#pragma unroll
  for (int k = 0; k < K; ++k) {
    u64 m = best[0];
#pragma unroll
    for (int off = G / 2; off > 0; off >>= 1) {
      u64 o = __shfl_xor_sync(gmask, m, off, G);
      m = o < m ? o : m;
    }
    res[k] = m;
    if (best[0] == m && m != kSentinel) {
#pragma unroll
      for (int s = 0; s < K - 1; ++s) best[s] = best[s + 1];
      best[K - 1] = kSentinel;
    }
  }
}

// Exact K-NN of (qx,qy,qz) by a cooperating group of G lanes (G = 8, 16 or 32, groups aligned inside a warp).
// Ring expansion over the voxel hash: the first pass visits the 3x3x3 block around the query voxel, further
// passes add one Chebyshev shell each.  The search stops when (a) the K-th distance is provably smaller than
// the distance to anything unvisited, (b) the visited block already contains the ball of radius sqrt(max_d2)
// (results beyond it are "don't care"), or (c) the block covers every occupied voxel.  After kMaxRing rings
// the group falls back to a coalesced brute-force sweep of the whole map.  res[] is identical in all lanes.
template <int K, int G>
__device__ __forceinline__ void knn_search(const GridView& g, float qx, float qy, float qz, float max_d2, unsigned lane,
                                           unsigned gmask, u64 (&res)[K]) {
  u64 best[K];
#pragma unroll
  for (int k = 0; k < K; ++k) best[k] = kSentinel, res[k] = kSentinel;
  if (g.n <= 0) return;

  float ux = __fmul_rn(qx, g.inv_cell), uy = __fmul_rn(qy, g.inv_cell), uz = __fmul_rn(qz, g.inv_cell);
  if (!(fabsf(ux) < (float)kCoordLim && fabsf(uy) < (float)kCoordLim && fabsf(uz) < (float)kCoordLim)) {
    // query outside the addressable voxel range (or NaN): brute force keeps the result exact.
    for (int j = lane; j < g.n; j += G) {
      float4 p = __ldg(g.sorted + j);
      topk_insert<K>(best, pack_cand(dist2_rn(qx, qy, qz, p.x, p.y, p.z), __float_as_uint(p.w)));
    }
    group_merge<K, G>(best, res, gmask);
    return;
  }
  int cx = __float2int_rd(ux), cy = __float2int_rd(uy), cz = __float2int_rd(uz);
  float fx = ux - (float)cx, fy = uy - (float)cy, fz = uz - (float)cz;
  float fmin = fminf(fminf(fminf(fx, 1.f - fx), fminf(fy, 1.f - fy)), fminf(fz, 1.f - fz));
  float umax = fmaxf(fmaxf(fabsf(ux), fabsf(uy)), fabsf(uz));
  int bx0 = __ldg(g.bbox + 0), by0 = __ldg(g.bbox + 1), bz0 = __ldg(g.bbox + 2);
  int bx1 = __ldg(g.bbox + 3), by1 = __ldg(g.bbox + 4), bz1 = __ldg(g.bbox + 5);

  int r = 1;
#pragma unroll 1
  for (;;) {
    const int side = 2 * r + 1, total = side * side * side;
#pragma unroll 1
    for (int t = lane; t < total; t += G) {
      int dz = t / (side * side), rem = t - dz * side * side;
      int dy = rem / side, dx = rem - dy * side;
      dx -= r, dy -= r, dz -= r;
      if (r > 1 && abs(dx) < r && abs(dy) < r && abs(dz) < r) continue;  // interior: visited by earlier passes
      int vx = cx + dx, vy = cy + dy, vz = cz + dz;
      if (vx < bx0 || vx > bx1 || vy < by0 || vy > by1 || vz < bz0 || vz > bz1) continue;
      scan_voxel<K>(g, vx, vy, vz, qx, qy, qz, best);
    }
    group_merge<K, G>(best, res, gmask);

    // distance (metres) from the query to the nearest unvisited voxel, made conservative against the float
    // rounding of the voxel coordinates of both the query and any map point (2^-24 relative each).
    float margin = (umax + (float)r + 2.f) * 2.4e-7f;
    float bound = ((float)r + fmin - margin) * g.cell;
    float b2 = bound > 0.f ? bound * bound * 0.999999f : 0.f;
    float dk = cand_d2(res[K - 1]);  // NaN pattern when fewer than K found
    bool done = (res[K - 1] != kSentinel && dk < b2) || (max_d2 > 0.f && b2 >= max_d2) ||
                (cx - r <= bx0 && cx + r >= bx1 && cy - r <= by0 && cy + r >= by1 && cz - r <= bz0 && cz + r >= bz1);
    if (done) return;
    if (r >= kMaxRing) break;
    ++r;
#pragma unroll
    for (int k = 0; k < K; ++k) best[k] = (lane == 0) ? res[k] : kSentinel;
  }
  // brute-force fallback (rare: sparse maps / far-away queries with no distance bound)
#pragma unroll
  for (int k = 0; k < K; ++k) best[k] = kSentinel;
  for (int j = lane; j < g.n; j += G) {
    float4 p = __ldg(g.sorted + j);
    topk_insert<K>(best, pack_cand(dist2_rn(qx, qy, qz, p.x, p.y, p.z), __float_as_uint(p.w)));
  }
  group_merge<K, G>(best, res, gmask);
}

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
