// ilsm_voxel.cuh -- block-level building blocks shared by the front end and the cube map: in-block bitonic sort
// and the PCL-style VoxelGrid of one block-resident point set.
#pragma once
#include "ilsm_internal.cuh"

namespace ilsm {

constexpr int kVoxelBlockMax = 16384;  // points one block can sort in shared memory (128 KB of u64 keys)

// In-block bitonic sort of P (power of two) unique 64-bit keys held in shared memory.
// Every thread owns E = P / blockDim.x consecutive keys in REGISTERS: compare-exchange steps whose partner distance j
// is below E stay inside the thread, steps with E <= j < 32 E go through warp shuffles, and only the steps with
// j >= 32 E (15 of the 91 steps at P = 8192 with 1024 threads) exchange through shared memory with a barrier.  The
// shared-memory exchange uses a striped layout (element e of thread t at e * blockDim + t): conflict-free.
// `base` = global index of keys[0] and [k_first, k_last] = the stages to run make the same routine one tile of a
// multi-block sort: a tile sorts itself with k = 2 .. P (direction from the GLOBAL index), and after the cross-tile
// exchanges of a stage k > P it finishes that stage's partner distances P/2 .. 1 (merge_only).
template <int E>
static __device__ __forceinline__ void bitonic_sort_regs(u64* keys, int P, unsigned base = 0u, unsigned k_first = 2u,
                                                         unsigned k_last = 0u, bool merge_only = false) {
  const int tid = threadIdx.x, nt = blockDim.x;
  if (k_last == 0u) k_last = (unsigned)P;
  u64 r[E];
#pragma unroll
  for (int e = 0; e < E; ++e) {
    const int idx = tid * E + e;
    r[e] = idx < P ? keys[idx] : ~0ull;
  }
  __syncthreads();
  const bool active = tid * E < P;  // P < blockDim.x: the idle threads hold sentinels and stay out of memory
  const int ns = E == 1 ? P : nt;   // stripe length (E > 1 implies P == E * blockDim.x)
  for (unsigned k = k_first; k <= k_last; k <<= 1) {
    int j = merge_only ? P >> 1 : (int)(k >> 1);
    for (; j >= 32 * E; j >>= 1) {  // partner in another warp: striped exchange through shared memory
      const int tj = j / E;
      if (active) {
#pragma unroll
        for (int e = 0; e < E; ++e) keys[e * ns + tid] = r[e];
      }
      __syncthreads();
      if (active) {
#pragma unroll
        for (int e = 0; e < E; ++e) {
          const int idx = tid * E + e;
          const u64 o = keys[e * ns + (tid ^ tj)];
          const bool want_min = ((idx & j) == 0) == (((base + (unsigned)idx) & k) == 0);
          r[e] = want_min ? (o < r[e] ? o : r[e]) : (o > r[e] ? o : r[e]);
        }
      }
      __syncthreads();
    }
    for (; j >= E; j >>= 1) {  // partner in another lane of this warp
      const int lj = j / E;
#pragma unroll
      for (int e = 0; e < E; ++e) {
        const int idx = tid * E + e;
        const u64 o = __shfl_xor_sync(0xffffffffu, r[e], lj);
        const bool want_min = ((idx & j) == 0) == (((base + (unsigned)idx) & k) == 0);
        r[e] = want_min ? (o < r[e] ? o : r[e]) : (o > r[e] ? o : r[e]);
      }
    }
    // partner inside the thread: compile-time distances keep r[] in registers
#pragma unroll
    for (int jj = E / 2; jj > 0; jj >>= 1) {
      if (jj <= j) {
#pragma unroll
        for (int e = 0; e < E; ++e) {
          if ((e & jj) == 0) {
            const bool up = ((base + (unsigned)(tid * E + e)) & k) == 0;
            const u64 a = r[e], b = r[e | jj];
            const bool sw = (a > b) == up;
            r[e] = sw ? b : a;
            r[e | jj] = sw ? a : b;
          }
        }
      }
    }
  }
#pragma unroll
  for (int e = 0; e < E; ++e) {
    const int idx = tid * E + e;
    if (idx < P) keys[idx] = r[e];
  }
  __syncthreads();
}

// P: power of two, at most 16 * blockDim.x; blockDim.x a power of two >= 32.  `keys` holds P entries.
static __device__ __forceinline__ void bitonic_sort_smem(u64* keys, int P) {
  const int per = (P + (int)blockDim.x - 1) / (int)blockDim.x;
  if (per <= 1) bitonic_sort_regs<1>(keys, P);
  else if (per == 2) bitonic_sort_regs<2>(keys, P);
  else if (per == 4) bitonic_sort_regs<4>(keys, P);
  else if (per == 8) bitonic_sort_regs<8>(keys, P);
  else bitonic_sort_regs<16>(keys, P);
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
      const u64 vox = keys[t] >> 24;
      int cnt = 1;  // run length first (shared memory only), so that the point loads below are independent of the sums
      while (t + cnt < P && (keys[t + cnt] >> 24) == vox) ++cnt;
      for (int e0 = 0; e0 < cnt; e0 += 4) {  // 4 gathers in flight, then the 4 sequential float adds (PCL's order)
        float4 q[4];
#pragma unroll
        for (int u = 0; u < 4; ++u)
          q[u] = e0 + u < cnt ? pts[(int)(keys[t + e0 + u] & 0xFFFFFF)] : make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
        for (int u = 0; u < 4; ++u)
          if (e0 + u < cnt)
            sx = __fadd_rn(sx, q[u].x), sy = __fadd_rn(sy, q[u].y), sz = __fadd_rn(sz, q[u].z), si = __fadd_rn(si, q[u].w);
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

// ---------------------------------------------------------------------------------------------------
// The same VoxelGrid without sorting the points: PCL orders its OUTPUT by voxel index and sums every voxel's points in
// input order, so only the distinct voxels have to be sorted (a stack of ~9 k less-flat points falls into ~2 k voxels of
// 0.8 m) and each voxel's members ranked by index.
//   1. bounding box -> voxel index per point (as above)
//   2. open-addressing hash of the voxel indices in shared memory (32768 slots): slot per point by atomicCAS, arrival
//      rank inside the voxel by atomicAdd (16-bit counters, two per word)
//   3. the occupied slots are compacted into a list (voxel index << 15 | slot; one atomic per warp) and sorted: ascending
//      voxel index
//   4. exclusive scan of the voxel populations in that order -> where each voxel's member list starts
//   5. every point drops its index into its voxel's list at its arrival rank (arbitrary order), then finds its place in
//      INDEX order by counting the smaller members of its list (all points in parallel, independent shared-memory
//      reads -- no per-voxel serial sort) and copies itself there: the points end up grouped by voxel, in input order
//   6. the float sums of pcl::CentroidPoint, bit for bit, over those contiguous runs: one thread per voxel for the short
//      runs, one warp per voxel for the long ones (32 coalesced loads in flight, the sequential adds fed by shuffles)
// Shared memory: smem must provide kVgHashSmemBytes (192 KB); scratch = kVgScratchWords 32-bit words of global memory per
// block.  m <= kVoxelBlockMax.  Returns the number of voxels (block-uniform).
// ---------------------------------------------------------------------------------------------------
constexpr int kVgHashSlots = 32768;
constexpr size_t kVgHashSmemBytes = (size_t)kVgHashSlots * 4 + (size_t)kVgHashSlots * 2;  // keys (reused: sort keys, member lists) + counters
constexpr size_t kVgScratchWords = (size_t)kVoxelBlockMax + kVgHashSlots + 2 * (size_t)kVoxelBlockMax + kVoxelBlockMax + 4 * (size_t)kVoxelBlockMax;
constexpr int kVgWarpRun = 16;  // runs longer than this are summed by a whole warp

static __device__ int voxelgrid_block_hash(const float4* __restrict__ pts, int m, float leaf, unsigned char* smem, uint32_t* __restrict__ scratch,
                                           float4* __restrict__ out, int* err) {
  __shared__ float s_min[3], s_max[3];
  __shared__ int s_V, s_carry, s_heavy;
  __shared__ int wc[32];
  uint32_t* hkeys = reinterpret_cast<uint32_t*>(smem);                       // region A, first life: the hash keys
  u64* skeys = reinterpret_cast<u64*>(smem);                                 // region A, second life: the voxel sort keys
  unsigned short* members = reinterpret_cast<unsigned short*>(smem);         // region A, third life: the member lists
  uint32_t* cnt32 = reinterpret_cast<uint32_t*>(smem + (size_t)kVgHashSlots * 4);  // region B: two 16-bit counters per word
  unsigned short* heavy = reinterpret_cast<unsigned short*>(cnt32);          // region B, second life: voxels with long runs
  uint32_t* g_sr = scratch;                                       // per point: slot << 16 | arrival rank (0xFFFFFFFF: skipped)
  uint32_t* g_sp = scratch + kVoxelBlockMax;                      // per slot: list start << 16 | output position
  u64* g_list = reinterpret_cast<u64*>(scratch + kVoxelBlockMax + kVgHashSlots);  // compacted (voxel << 15 | slot)
  uint32_t* g_vinfo = scratch + kVoxelBlockMax + kVgHashSlots + 2 * kVoxelBlockMax;  // per output voxel: start << 16 | count
  float4* g_run = reinterpret_cast<float4*>(scratch + kVoxelBlockMax + kVgHashSlots + 3 * (size_t)kVoxelBlockMax);  // points by (voxel, index)
  const int tid = threadIdx.x, nt = blockDim.x;
  const int lane = tid & 31, warp = tid >> 5, nw = nt >> 5;
  if (tid < 3) s_min[tid] = __int_as_float(0x7f800000), s_max[tid] = __int_as_float(0xff800000);
  if (tid == 0) s_V = 0, s_carry = 0, s_heavy = 0;
  for (int i = tid; i < kVgHashSlots; i += nt) hkeys[i] = 0xFFFFFFFFu;
  for (int i = tid; i < kVgHashSlots / 2; i += nt) cnt32[i] = 0u;
  __syncthreads();
  float mn[3] = {__int_as_float(0x7f800000), __int_as_float(0x7f800000), __int_as_float(0x7f800000)};
  float mx[3] = {__int_as_float(0xff800000), __int_as_float(0xff800000), __int_as_float(0xff800000)};
  for (int t = tid; t < m; t += nt) {
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
  if (lane == 0) {
#pragma unroll
    for (int a = 0; a < 3; ++a) {
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
    if (tid == 0) atomicOr(err, 2);
    return 0;
  }
  // ---- 2. voxel slot and arrival rank of every point
  for (int t = tid; t < m; t += nt) {
    const float4 p = pts[t];
    uint32_t sr = 0xFFFFFFFFu;
    if (isfinite(p.x) && isfinite(p.y) && isfinite(p.z)) {
      const int i0 = __float2int_rz(__fsub_rn(floorf(__fmul_rn(p.x, inv)), (float)min_b[0]));
      const int i1 = __float2int_rz(__fsub_rn(floorf(__fmul_rn(p.y, inv)), (float)min_b[1]));
      const int i2 = __float2int_rz(__fsub_rn(floorf(__fmul_rn(p.z, inv)), (float)min_b[2]));
      const uint32_t idx = (uint32_t)(i0 + i1 * mul1 + i2 * mul2);
      uint32_t slot = (idx * 2654435761u) >> 17;
      for (;;) {
        const uint32_t prev = atomicCAS(&hkeys[slot], 0xFFFFFFFFu, idx);
        if (prev == 0xFFFFFFFFu || prev == idx) break;
        slot = (slot + 1) & (kVgHashSlots - 1);
      }
      const int sh = 16 * (slot & 1);
      const uint32_t old = atomicAdd(&cnt32[slot >> 1], 1u << sh);
      sr = (slot << 16) | ((old >> sh) & 0xFFFFu);
    }
    g_sr[t] = sr;
  }
  __syncthreads();
  // ---- 3. occupied slots -> list (one shared-memory atomic per warp and pass), sorted by voxel index
  for (int sl0 = 0; sl0 < kVgHashSlots; sl0 += nt) {
    const int sl = sl0 + tid;
    const uint32_t key = hkeys[sl];
    const bool occ = key != 0xFFFFFFFFu;
    const unsigned b = __ballot_sync(0xffffffffu, occ);
    int wbase = 0;
    if (lane == 0 && b) wbase = atomicAdd(&s_V, __popc(b));
    wbase = __shfl_sync(0xffffffffu, wbase, 0);
    if (occ) g_list[wbase + __popc(b & ((1u << lane) - 1u))] = ((u64)key << 15) | (u64)sl;
  }
  __syncthreads();
  const int V = s_V;
  int Pv = 32;
  while (Pv < V) Pv <<= 1;
  for (int j = tid; j < Pv; j += nt) skeys[j] = j < V ? g_list[j] : ~0ull;  // region A changes hands: every thread is past its hkeys reads
  __syncthreads();
  bitonic_sort_smem(skeys, Pv);
  // ---- 4. exclusive scan of the populations in output order
  for (int j0 = 0; j0 < Pv; j0 += nt) {
    const int j = j0 + tid;
    uint32_t slot = 0;
    int c = 0;
    if (j < V) {
      slot = (uint32_t)(skeys[j] & 0x7FFFu);
      c = (int)((cnt32[slot >> 1] >> (16 * (slot & 1))) & 0xFFFFu);
    }
    int inc = c;
#pragma unroll
    for (int off = 1; off < 32; off <<= 1) {
      const int o = __shfl_up_sync(0xffffffffu, inc, off);
      if (lane >= off) inc += o;
    }
    if (lane == 31) wc[warp] = inc;
    __syncthreads();
    int wbase = 0, tot = 0;
    for (int w = 0; w < nw; ++w) {
      if (w < warp) wbase += wc[w];
      tot += wc[w];
    }
    const int start = s_carry + wbase + inc - c;
    if (j < V) {
      g_sp[slot] = ((uint32_t)start << 16) | (uint32_t)j;
      g_vinfo[j] = ((uint32_t)start << 16) | (uint32_t)c;
    }
    __syncthreads();
    if (tid == 0) s_carry += tot;
    __syncthreads();
  }
  __threadfence_block();
  // ---- 5. member lists (region A changes hands again: the sort keys are consumed) ...
  for (int t = tid; t < m; t += nt) {
    const uint32_t sr = g_sr[t];
    if (sr != 0xFFFFFFFFu) members[(g_sp[sr >> 16] >> 16) + (sr & 0xFFFFu)] = (unsigned short)t;
  }
  __syncthreads();
  // ... and every point to its place in (voxel, index) order; the long runs are queued for the warps (region B is free)
  for (int t = tid; t < m; t += nt) {
    const uint32_t sr = g_sr[t];
    if (sr == 0xFFFFFFFFu) continue;
    const uint32_t sp = g_sp[sr >> 16];
    const int start = (int)(sp >> 16), c = (int)(g_vinfo[sp & 0xFFFFu] & 0xFFFFu);
    int rank = 0;
    if (c > 1) {
      const unsigned short* L = members + start;
      int u = 0;
      if ((start & 1) && c > 0) rank += L[0] < t, u = 1;  // align to a 32-bit word, then two members per load
      for (; u + 1 < c; u += 2) {
        const uint32_t w2 = *reinterpret_cast<const uint32_t*>(L + u);
        rank += (int)((w2 & 0xFFFFu) < (uint32_t)t) + (int)((w2 >> 16) < (uint32_t)t);
      }
      if (u < c) rank += L[u] < t;
    }
    g_run[start + rank] = pts[t];
  }
  for (int j = tid; j < V; j += nt)
    if ((int)(g_vinfo[j] & 0xFFFFu) > kVgWarpRun) heavy[atomicAdd(&s_heavy, 1)] = (unsigned short)j;
  __syncthreads();
  // ---- 6. the ordered float sums: short runs, one thread per voxel
  for (int j = tid; j < V; j += nt) {
    const uint32_t vi = g_vinfo[j];
    const int start = (int)(vi >> 16), c = (int)(vi & 0xFFFFu);
    if (c > kVgWarpRun) continue;
    const float4* R = g_run + start;
    float sx = 0.f, sy = 0.f, sz = 0.f, si = 0.f;
    for (int e0 = 0; e0 < c; e0 += 8) {  // 8 loads in flight, then the sequential float adds (PCL's order)
      float4 q[8];
#pragma unroll
      for (int k = 0; k < 8; ++k) q[k] = e0 + k < c ? R[e0 + k] : make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
      for (int k = 0; k < 8; ++k)
        if (e0 + k < c) sx = __fadd_rn(sx, q[k].x), sy = __fadd_rn(sy, q[k].y), sz = __fadd_rn(sz, q[k].z), si = __fadd_rn(si, q[k].w);
    }
    const float cf = (float)c;
    out[j] = make_float4(__fdiv_rn(sx, cf), __fdiv_rn(sy, cf), __fdiv_rn(sz, cf), __fdiv_rn(si, cf));
  }
  // long runs, one warp per voxel: 32 coalesced loads per round (the next round's already in flight), every lane adds
  // the members up in order from shuffles
  const int H = s_heavy;
  for (int h = warp; h < H; h += nw) {
    const int j = heavy[h];
    const uint32_t vi = g_vinfo[j];
    const int start = (int)(vi >> 16), c = (int)(vi & 0xFFFFu);
    const float4* R = g_run + start;
    float sx = 0.f, sy = 0.f, sz = 0.f, si = 0.f;
    float4 nxt = lane < c ? R[lane] : make_float4(0.f, 0.f, 0.f, 0.f);
    for (int e0 = 0; e0 < c; e0 += 32) {
      const float4 q = nxt;
      if (e0 + 32 + lane < c) nxt = R[e0 + 32 + lane];
      const int lim = c - e0 < 32 ? c - e0 : 32;
      for (int l = 0; l < lim; ++l) {
        sx = __fadd_rn(sx, __shfl_sync(0xffffffffu, q.x, l)), sy = __fadd_rn(sy, __shfl_sync(0xffffffffu, q.y, l));
        sz = __fadd_rn(sz, __shfl_sync(0xffffffffu, q.z, l)), si = __fadd_rn(si, __shfl_sync(0xffffffffu, q.w, l));
      }
    }
    if (lane == 0) {
      const float cf = (float)c;
      out[j] = make_float4(__fdiv_rn(sx, cf), __fdiv_rn(sy, cf), __fdiv_rn(sz, cf), __fdiv_rn(si, cf));
    }
  }
  __syncthreads();
  return V;
}

// ---------------------------------------------------------------------------------------------------
// VoxelGrid of a cloud whose first n_old points are the OUTPUT of an earlier VoxelGrid pass at the same leaf size (a
// map cube between two frames: laserMapping.cpp:987-1002 re-filters it after every insertion) followed by n - n_old new
// points.  PCL's result is independent per voxel and ordered by voxel index, and the relative order of two voxel
// indices does not depend on the bounding box, so when the old points still sit one per voxel in ascending voxel order
// (checked -- a centroid that rounded across a voxel face sends the block to the general path) the pass is a MERGE:
//   * only the new points are sorted, by (voxel, index);
//   * every distinct new voxel finds its old centroid by binary search -- present: that centroid (lowest index, summed
//     first) plus the new members; absent: a new output voxel;
//   * every untouched old point moves up by the number of new voxels below it (one binary search), its value passing
//     through the same float operations as a one-member voxel, (0 + p) / 1.
// Bit-identical to the general path by construction; O(n log m) instead of a hash + sort of everything, and one pass over
// the cube instead of three (the keys are taken relative to the first point's voxel: no bounding box is needed for an order).
// Shared memory: 64 KB of old keys + 16 KB of new sort keys + 20 KB of per-new-voxel tables.  Returns the number of
// voxels, or -1 (block-uniform) when the preconditions do not hold.
// ---------------------------------------------------------------------------------------------------
constexpr int kVgMergeNew = 2048;

static __device__ __forceinline__ int block_excl_scan_1024(int v, int* wc, int* total) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nw = blockDim.x >> 5;
  int inc = v;
#pragma unroll
  for (int off = 1; off < 32; off <<= 1) {
    const int o = __shfl_up_sync(0xffffffffu, inc, off);
    if (lane >= off) inc += o;
  }
  __syncthreads();  // wc may still be read from the previous call
  if (lane == 31) wc[warp] = inc;
  __syncthreads();
  int wbase = 0, tot = 0;
  for (int w = 0; w < nw; ++w) {
    if (w < warp) wbase += wc[w];
    tot += wc[w];
  }
  *total = tot;
  return wbase + inc - v;
}

static __device__ int voxelgrid_block_merge(const float4* __restrict__ pts, int n_old, int n, float leaf, unsigned char* smem,
                                            float4* __restrict__ out, int* err) {
  __shared__ int s_ref[3];
  __shared__ int wc[32];
  uint32_t* ek = reinterpret_cast<uint32_t*>(smem);                                     // old keys (ascending when valid)
  u64* nk = reinterpret_cast<u64*>(smem + (size_t)kVoxelBlockMax * 4);                  // new (voxel << 24 | position)
  uint32_t* dkey = reinterpret_cast<uint32_t*>(smem + (size_t)kVoxelBlockMax * 4 + (size_t)kVgMergeNew * 8);  // distinct new voxels
  unsigned short* dstart = reinterpret_cast<unsigned short*>(dkey + kVgMergeNew);       // first member of each in nk (+ end)
  unsigned short* dexcl = dstart + kVgMergeNew + 2;                                     // new-only voxels before each (+ total)
  const int tid = threadIdx.x, nt = blockDim.x;
  const int m = n - n_old;
  // Voxel keys RELATIVE TO THE FIRST POINT'S VOXEL instead of the bounding box PCL uses: only the order of two keys and their
  // equality matter here, and (z, y, x)-lexicographic order does not depend on the origin -- this saves the bounding-box pass
  // over the cube.  A cube spans 50 m (125 voxels of 0.4 m), so +-255 voxels per axis is ample; anything outside (or a
  // non-finite point) sends the block to the general path, which also owns PCL's "leaf size too small" diagnosis.
  const float inv = __fdiv_rn(1.0f, leaf);
  if (tid == 0) {
    const float4 p = pts[0];
    s_ref[0] = __float2int_rz(floorf(__fmul_rn(p.x, inv))), s_ref[1] = __float2int_rz(floorf(__fmul_rn(p.y, inv)));
    s_ref[2] = __float2int_rz(floorf(__fmul_rn(p.z, inv)));
  }
  __syncthreads();
  const int r0 = s_ref[0], r1 = s_ref[1], r2 = s_ref[2];
  int bad = 0;
  int P = 32;
  while (P < m) P <<= 1;
  for (int t = tid; t < n_old + P; t += nt) {
    uint32_t idx = 0xFFFFFFFFu;
    if (t < n) {
      const float4 p = pts[t];
      if (isfinite(p.x) && isfinite(p.y) && isfinite(p.z)) {
        const int i0 = __float2int_rz(floorf(__fmul_rn(p.x, inv))) - r0 + 256, i1 = __float2int_rz(floorf(__fmul_rn(p.y, inv))) - r1 + 256;
        const int i2 = __float2int_rz(floorf(__fmul_rn(p.z, inv))) - r2 + 256;
        if ((unsigned)i0 < 512u && (unsigned)i1 < 512u && (unsigned)i2 < 512u) idx = (uint32_t)(i0 | (i1 << 9) | (i2 << 18));
        else bad = 1;
      } else {
        bad = 1;
      }
    }
    if (t < n_old) ek[t] = idx;
    else nk[t - n_old] = t < n ? (((u64)idx << 24) | (u64)(t - n_old)) : ~0ull;
  }
  __syncthreads();
  for (int t = tid + 1; t < n_old; t += nt) bad |= ek[t - 1] >= ek[t];
  if (__syncthreads_or(bad)) return -1;
  bitonic_sort_smem(nk, P);
  // distinct new voxels
  int D = 0;
  for (int i0 = 0; i0 < m; i0 += nt) {
    const int i = i0 + tid;
    const bool head = i < m && (i == 0 || (nk[i] >> 24) != (nk[i - 1] >> 24));
    int tot;
    const int d = D + block_excl_scan_1024(head ? 1 : 0, wc, &tot);
    if (head) dkey[d] = (uint32_t)(nk[i] >> 24), dstart[d] = (unsigned short)i;
    D += tot;
  }
  if (tid == 0) dstart[D] = (unsigned short)m;
  __syncthreads();
  // each distinct new voxel against the old keys; new-only voxels open an output slot
  int n_newonly = 0;
  for (int d0 = 0; d0 < D; d0 += nt) {
    const int d = d0 + tid;
    int pos = 0;
    bool exists = false;
    uint32_t key = 0;
    if (d < D) {
      key = dkey[d];
      int lo = 0, hi = n_old;
      while (lo < hi) {
        const int mid = (lo + hi) >> 1;
        if (ek[mid] < key) lo = mid + 1; else hi = mid;
      }
      pos = lo;
      exists = pos < n_old && ek[pos] == key;
    }
    int tot;
    const int before = n_newonly + block_excl_scan_1024(d < D && !exists ? 1 : 0, wc, &tot);
    if (d < D) {
      dexcl[d] = (unsigned short)before;
      const int b = dstart[d], c = (int)dstart[d + 1] - b;
      float sx = 0.f, sy = 0.f, sz = 0.f, si = 0.f;
      if (exists) {
        const float4 e = pts[pos];
        sx = __fadd_rn(sx, e.x), sy = __fadd_rn(sy, e.y), sz = __fadd_rn(sz, e.z), si = __fadd_rn(si, e.w);
      }
      for (int e0 = 0; e0 < c; e0 += 4) {
        float4 q[4];
#pragma unroll
        for (int k = 0; k < 4; ++k)
          q[k] = e0 + k < c ? pts[n_old + (int)(nk[b + e0 + k] & 0xFFFFFFull)] : make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
        for (int k = 0; k < 4; ++k)
          if (e0 + k < c) sx = __fadd_rn(sx, q[k].x), sy = __fadd_rn(sy, q[k].y), sz = __fadd_rn(sz, q[k].z), si = __fadd_rn(si, q[k].w);
      }
      const float cf = (float)(c + (exists ? 1 : 0));
      out[pos + before] = make_float4(__fdiv_rn(sx, cf), __fdiv_rn(sy, cf), __fdiv_rn(sz, cf), __fdiv_rn(si, cf));
    }
    n_newonly += tot;
  }
  if (tid == 0) dexcl[D] = (unsigned short)n_newonly;
  __syncthreads();
  // the untouched old points
  for (int t = tid; t < n_old; t += nt) {
    const uint32_t key = ek[t];
    int lo = 0, hi = D;
    while (lo < hi) {
      const int mid = (lo + hi) >> 1;
      if (dkey[mid] < key) lo = mid + 1; else hi = mid;
    }
    if (lo < D && dkey[lo] == key) continue;  // summed with its new members above
    const float4 p = pts[t];
    out[t + dexcl[lo]] = make_float4(__fdiv_rn(__fadd_rn(0.f, p.x), 1.f), __fdiv_rn(__fadd_rn(0.f, p.y), 1.f),
                                     __fdiv_rn(__fadd_rn(0.f, p.z), 1.f), __fdiv_rn(__fadd_rn(0.f, p.w), 1.f));
  }
  __syncthreads();
  return n_old + n_newonly;
}

}  // namespace ilsm
