// ilsm_voxel.cuh -- block-level building blocks shared by the front end and the cube map: in-block bitonic sort
// and the PCL-style VoxelGrid of one block-resident point set.
#pragma once
#include "ilsm_internal.cuh"

namespace ilsm {

constexpr int kVoxelBlockMax = 16384;  // points one block can sort in shared memory (128 KB of u64 keys)

// in-block bitonic sort of P (power of two) 64-bit keys in shared memory
static __device__ __forceinline__ void bitonic_sort_smem(u64* keys, int P) {
  for (int k = 2; k <= P; k <<= 1) {
    for (int j = k >> 1; j > 0; j >>= 1) {
      for (int t = threadIdx.x; t < P; t += blockDim.x) {
        const int ixj = t ^ j;
        if (ixj > t) {
          const u64 a = keys[t], b = keys[ixj];
          const bool up = (t & k) == 0;
          if ((a > b) == up) {
            keys[t] = b;
            keys[ixj] = a;
          }
        }
      }
      __syncthreads();
    }
  }
}

// PCL 1.10 VoxelGrid over one contiguous set of points held by the block (PCL 1.10 applyFilter restated):
// bounding box -> voxel index -> stable sort by (voxel, point order) -> float centroid per voxel in that order.
// pts: m points (global), out: centroids in ascending voxel index; returns the number of voxels (block-uniform).
// keys: shared scratch of P >= m (power of two) entries.
static __device__ int voxelgrid_block(const float4* __restrict__ pts, int m, float leaf, u64* keys, int P, float4* __restrict__ out,
                               int* err) {
  __shared__ float s_min[3], s_max[3];
  __shared__ int s_count, s_bad;
  if (threadIdx.x < 3) s_min[threadIdx.x] = __int_as_float(0x7f800000), s_max[threadIdx.x] = __int_as_float(0xff800000);
  if (threadIdx.x == 0) s_count = 0, s_bad = 0;
  __syncthreads();
  float mn[3] = {__int_as_float(0x7f800000), __int_as_float(0x7f800000), __int_as_float(0x7f800000)};
  float mx[3] = {__int_as_float(0xff800000), __int_as_float(0xff800000), __int_as_float(0xff800000)};
  for (int t = threadIdx.x; t < m; t += blockDim.x) {
    const float4 p = pts[t];
    if (isfinite(p.x) && isfinite(p.y) && isfinite(p.z)) {
      mn[0] = fminf(mn[0], p.x), mn[1] = fminf(mn[1], p.y), mn[2] = fminf(mn[2], p.z);
      mx[0] = fmaxf(mx[0], p.x), mx[1] = fmaxf(mx[1], p.y), mx[2] = fmaxf(mx[2], p.z);
    }
  }
#pragma unroll
  for (int a = 0; a < 3; ++a) {
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) {
      mn[a] = fminf(mn[a], __shfl_xor_sync(0xffffffffu, mn[a], off));
      mx[a] = fmaxf(mx[a], __shfl_xor_sync(0xffffffffu, mx[a], off));
    }
  }
  if ((threadIdx.x & 31) == 0) {
#pragma unroll
    for (int a = 0; a < 3; ++a) {
      // float atomic min/max through the ordered-int trick (values finite or +-inf)
      int* pmn = reinterpret_cast<int*>(&s_min[a]);
      int* pmx = reinterpret_cast<int*>(&s_max[a]);
      if (mn[a] >= 0.f) atomicMin(pmn, __float_as_int(mn[a])); else atomicMax(reinterpret_cast<unsigned*>(pmn), __float_as_uint(mn[a]));
      if (mx[a] >= 0.f) atomicMax(pmx, __float_as_int(mx[a])); else atomicMin(reinterpret_cast<unsigned*>(pmx), __float_as_uint(mx[a]));
    }
  }
  __syncthreads();
  const float inv = __fdiv_rn(1.0f, leaf);
  int min_b[3], div_b[3];
#pragma unroll
  for (int a = 0; a < 3; ++a) {
    min_b[a] = __float2int_rd(__fmul_rn(s_min[a], inv));
    div_b[a] = __float2int_rd(__fmul_rn(s_max[a], inv)) - min_b[a] + 1;
  }
  const long long mul1 = div_b[0], mul2 = (long long)div_b[0] * div_b[1];
  if (mul2 * div_b[2] >= (1ll << 31)) {  // pcl: "Leaf size is too small for the input dataset"
    if (threadIdx.x == 0) atomicOr(err, 2);
    return 0;
  }
  for (int t = threadIdx.x; t < P; t += blockDim.x) {
    u64 key = ~0ull;
    if (t < m) {
      const float4 p = pts[t];
      if (isfinite(p.x) && isfinite(p.y) && isfinite(p.z)) {
        const int i0 = __float2int_rz(__fsub_rn(floorf(__fmul_rn(p.x, inv)), (float)min_b[0]));
        const int i1 = __float2int_rz(__fsub_rn(floorf(__fmul_rn(p.y, inv)), (float)min_b[1]));
        const int i2 = __float2int_rz(__fsub_rn(floorf(__fmul_rn(p.z, inv)), (float)min_b[2]));
        const long long idx = i0 + i1 * mul1 + i2 * mul2;
        key = ((u64)idx << 24) | (uint32_t)t;  // (voxel, point order): stable by construction
      }
    }
    keys[t] = key;
  }
  __syncthreads();
  bitonic_sort_smem(keys, P);
  // run heads -> output slot = number of heads before; each head accumulates its run in order (float, like
  // pcl::CentroidPoint) and divides by the count
  for (int t0 = 0; t0 < P; t0 += blockDim.x) {
    const int t = t0 + threadIdx.x;
    bool head = false;
    if (t < P && keys[t] != ~0ull) head = t == 0 || (keys[t] >> 24) != (keys[t - 1] >> 24);
    // block-wide exclusive count of heads in this chunk
    const unsigned b = __ballot_sync(0xffffffffu, head);
    __shared__ int wc[32];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    if (lane == 0) wc[warp] = __popc(b);
    __syncthreads();
    int wbase = 0, tot = 0;
    const int nw = blockDim.x >> 5;
    for (int w = 0; w < nw; ++w) {
      if (w < warp) wbase += wc[w];
      tot += wc[w];
    }
    const int base = s_count;
    if (head) {
      const int slot = base + wbase + __popc(b & ((1u << lane) - 1u));
      float sx = 0.f, sy = 0.f, sz = 0.f, si = 0.f;
      int cnt = 0;
      const u64 vox = keys[t] >> 24;
      for (int e = t; e < P && keys[e] != ~0ull && (keys[e] >> 24) == vox; ++e) {
        const float4 p = pts[(int)(keys[e] & 0xFFFFFF)];
        sx = __fadd_rn(sx, p.x), sy = __fadd_rn(sy, p.y), sz = __fadd_rn(sz, p.z), si = __fadd_rn(si, p.w);
        ++cnt;
      }
      const float c = (float)cnt;
      out[slot] = make_float4(__fdiv_rn(sx, c), __fdiv_rn(sy, c), __fdiv_rn(sz, c), __fdiv_rn(si, c));
    }
    __syncthreads();
    if (threadIdx.x == 0) s_count = base + tot;
    __syncthreads();
  }
  return s_count;
}


}  // namespace ilsm
