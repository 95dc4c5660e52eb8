// scancontext.cu -- K5: ScanContext descriptor (makeScancontext) and loop-closure candidate scoring over a
// (sharded) keyframe database with local top-k.
//
// Replaces SCManager::makeScancontext (Scancontext.cpp:160-204), makeSectorkeyFromScancontext (:222-235),
// fastAlignUsingVkey (:104-124), distDirectSC (:79-101), distanceBtnScanContext (:126-157) and the candidate loop
// of detectLoopClosureID (:299-312), scored over EVERY database entry instead of the 10 ring-key candidates (a
// superset of the reference's search; each rank of a sharded run keeps its local top-k, the ranks exchange
// k x 16 B with one all-gather and merge identically).
//
// Descriptor values are floats in the reference (SCPointType is float, widened into a MatrixXd), so the database
// is stored as float32 [20][60] row-major = 4800 B per keyframe without loss; all arithmetic is fp64 like Eigen's.
// The kernel is HBM-bound for one query (4800 B per candidate, ~12 kFLOP fp64); no tensor cores: the 7-shift
// column-cosine search after sector-key alignment is not a dense contraction at batch size 1.
#include "ilsm_host.hpp"

namespace ilsm {

constexpr int kNR = 20, kNS = 60, kDesc = kNR * kNS;

struct ScQuery {  // prepared once per query, read by every scoring block
  double desc[kDesc];
  double norm[kNS];
  double key[kNS];
};

// ---------------------------------------------------------------------------------------------------
// makeScancontext: max (z + 2.0) per polar bin, empty bins -> 0
// ---------------------------------------------------------------------------------------------------
__device__ __forceinline__ float atan_rf(float v) { return (float)atan((double)v); }
__device__ __forceinline__ float xy2theta_dev(float x, float y) {
  const double k = 180.0 / 3.14159265358979323846;
  if (x >= 0 && y >= 0) return (float)(k * (double)atan_rf(__fdiv_rn(y, x)));
  if (x < 0 && y >= 0) return (float)(180.0 - (k * (double)atan_rf(__fdiv_rn(y, -x))));
  if (x < 0 && y < 0) return (float)(180.0 + (k * (double)atan_rf(__fdiv_rn(y, x))));
  return (float)(360.0 - (k * (double)atan_rf(__fdiv_rn(-y, x))));
}

__device__ __forceinline__ int float_order(float f) {  // monotone float -> int map for atomicMax
  const int i = __float_as_int(f);
  return i >= 0 ? i : i ^ 0x7FFFFFFF;
}
__device__ __forceinline__ float order_float(int i) { return __int_as_float(i >= 0 ? i : i ^ 0x7FFFFFFF); }

__global__ void sc_make_clear_kernel(int* bins) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < kDesc) bins[i] = float_order(-1000.0f);
}
__global__ void sc_make_kernel(const float* __restrict__ pts, int n, int stride_f, int* bins) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const float* p = pts + (size_t)i * stride_f;
  const float x = __ldg(p), y = __ldg(p + 1);
  const float z = (float)((double)__ldg(p + 2) + 2.0);  // LIDAR_HEIGHT, Scancontext.h:77
  const float range = __fsqrt_rn(__fadd_rn(__fmul_rn(x, x), __fmul_rn(y, y)));
  const float angle = xy2theta_dev(x, y);
  if ((double)range > 80.0) return;
  int ring = __double2int_ru(((double)range / 80.0) * 20.0);
  int sector = __double2int_ru(((double)angle / 360.0) * 60.0);
  ring = max(min(kNR, ring), 1);
  sector = max(min(kNS, sector), 1);
  atomicMax(&bins[(ring - 1) * kNS + (sector - 1)], float_order(z));
}
__global__ void sc_make_finish_kernel(const int* bins, float* desc) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < kDesc) {
    const float v = order_float(bins[i]);
    desc[i] = v == -1000.0f ? 0.f : v;
  }
}

// ---------------------------------------------------------------------------------------------------
// scoring
// ---------------------------------------------------------------------------------------------------
__global__ void sc_query_prep_kernel(const float* __restrict__ qdesc, ScQuery* q) {
  const int c = threadIdx.x;
  for (int i = threadIdx.x; i < kDesc; i += blockDim.x) q->desc[i] = (double)qdesc[i];
  if (c < kNS) {
    double s = 0, n2 = 0;
    for (int r = 0; r < kNR; ++r) {
      const double v = (double)qdesc[r * kNS + c];
      s += v;
      n2 += v * v;
    }
    q->key[c] = s / kNR;
    q->norm[c] = sqrt(n2);
  }
}

constexpr int kScWarps = 4;

// one warp per candidate keyframe, grid-stride
__global__ void __launch_bounds__(kScWarps * 32)
    sc_score_kernel(const float* __restrict__ db, int n, const ScQuery* __restrict__ q, double* __restrict__ out_dist,
                    int* __restrict__ out_shift) {
  __shared__ double qd[kDesc];
  __shared__ double qn[kNS], qk[kNS];
  __shared__ float cd[kScWarps][kDesc];
  __shared__ double ck[kScWarps][kNS], cn[kScWarps][kNS];
  for (int i = threadIdx.x; i < kDesc; i += blockDim.x) qd[i] = q->desc[i];
  if (threadIdx.x < kNS) qn[threadIdx.x] = q->norm[threadIdx.x], qk[threadIdx.x] = q->key[threadIdx.x];
  __syncthreads();
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int nwarps = gridDim.x * kScWarps;
  for (int cand = blockIdx.x * kScWarps + warp; cand < n; cand += nwarps) {
    // stage the candidate (4800 B, coalesced 128-bit loads)
    const float4* src = reinterpret_cast<const float4*>(db + (size_t)cand * kDesc);
    float4* dst = reinterpret_cast<float4*>(cd[warp]);
    for (int i = lane; i < kDesc / 4; i += 32) dst[i] = __ldg(src + i);
    __syncwarp();
    // sector key (column means) and column norms of the candidate
    for (int c = lane; c < kNS; c += 32) {
      double s = 0, n2 = 0;
#pragma unroll 4
      for (int r = 0; r < kNR; ++r) {
        const double v = (double)cd[warp][r * kNS + c];
        s += v;
        n2 += v * v;
      }
      ck[warp][c] = s / kNR;
      cn[warp][c] = sqrt(n2);
    }
    __syncwarp();
    // fastAlignUsingVkey: argmin_s || qk - circshift(ck, s) ||, first minimum wins
    double best = 10000000;
    int arg = 0;
    for (int s = lane; s < kNS; s += 32) {
      double ss = 0;
      for (int c = 0; c < kNS; ++c) {
        int cb = c - s;
        cb += cb < 0 ? kNS : 0;
        const double d = qk[c] - ck[warp][cb];
        ss += d * d;
      }
      const double nrm = sqrt(ss);
      if (nrm < best) best = nrm, arg = s;  // ascending s within the lane
    }
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) {
      const double ob = __shfl_xor_sync(0xffffffffu, best, off);
      const int oa = __shfl_xor_sync(0xffffffffu, arg, off);
      if (ob < best || (ob == best && oa < arg)) best = ob, arg = oa;
    }
    // distDirectSC on the 7 shifts around the aligned one, ascending shift order, first minimum wins
    int shifts[7];
#pragma unroll
    for (int k = 0; k < 7; ++k) shifts[k] = (arg + (k - 3) + kNS) % kNS;
    // sort ascending (7 values, a rotation: at most one wrap point)
#pragma unroll
    for (int a = 0; a < 6; ++a)
#pragma unroll
      for (int b = 0; b < 6 - a; ++b)
        if (shifts[b] > shifts[b + 1]) {
          const int t = shifts[b];
          shifts[b] = shifts[b + 1];
          shifts[b + 1] = t;
        }
    double sum[7];
    int eff[7];
#pragma unroll
    for (int k = 0; k < 7; ++k) sum[k] = 0.0, eff[k] = 0;
    for (int j = lane; j < kNS; j += 32) {
      const double nq = qn[j];
#pragma unroll
      for (int k = 0; k < 7; ++k) {
        int jb = j - shifts[k];
        jb += jb < 0 ? kNS : 0;
        const double nc = cn[warp][jb];
        if (nq != 0.0 && nc != 0.0) {
          double dot = 0;
#pragma unroll 4
          for (int r = 0; r < kNR; ++r) dot += qd[r * kNS + j] * (double)cd[warp][r * kNS + jb];
          sum[k] += dot / (nq * nc);
          eff[k] += 1;
        }
      }
    }
    double bd = 10000000;  // min_sc_dist / argmin_shift initial values of distanceBtnScanContext
    int bs = 0;
#pragma unroll
    for (int k = 0; k < 7; ++k) {
      double sk = sum[k];
      int ek = eff[k];
#pragma unroll
      for (int off = 16; off > 0; off >>= 1) {
        sk += __shfl_xor_sync(0xffffffffu, sk, off);
        ek += __shfl_xor_sync(0xffffffffu, ek, off);
      }
      const double d = ek > 0 ? 1.0 - sk / (double)ek : __longlong_as_double(0x7ff0000000000000ll);
      if (d < bd) bd = d, bs = shifts[k];
    }
    if (lane == 0) {
      out_dist[cand] = bd;
      out_shift[cand] = bs;
    }
    __syncwarp();
  }
}

// top-k of (dist, id) over n scored candidates by one block; k <= 16
constexpr int kTopKMax = 16;
struct ScKey {
  u64 d;  // distance bits (>= 0 => monotone as unsigned)
  int id;
};
__device__ __forceinline__ bool sc_less(const ScKey& a, const ScKey& b) { return a.d < b.d || (a.d == b.d && a.id < b.id); }

__global__ void __launch_bounds__(1024) sc_topk_kernel(const double* __restrict__ dist, const int* __restrict__ shift, int n,
                                                       int id_offset, int k, double* __restrict__ o_dist,
                                                       int* __restrict__ o_id, int* __restrict__ o_shift) {
  __shared__ u64 s_d[32];
  __shared__ int s_id[32];
  __shared__ int s_win;
  ScKey best[kTopKMax];
#pragma unroll
  for (int j = 0; j < kTopKMax; ++j) best[j].d = ~0ull, best[j].id = INT_MAX;
  for (int i = threadIdx.x; i < n; i += blockDim.x) {
    ScKey c;
    c.d = (u64)__double_as_longlong(dist[i]);
    c.id = i;
    if (sc_less(c, best[kTopKMax - 1])) {
      best[kTopKMax - 1] = c;
#pragma unroll
      for (int s = kTopKMax - 1; s > 0; --s) {
        if (sc_less(best[s], best[s - 1])) {
          const ScKey t = best[s];
          best[s] = best[s - 1];
          best[s - 1] = t;
        }
      }
    }
  }
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  for (int round = 0; round < k; ++round) {
    ScKey m = best[0];
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) {
      ScKey o;
      o.d = __shfl_xor_sync(0xffffffffu, m.d, off);
      o.id = __shfl_xor_sync(0xffffffffu, m.id, off);
      if (sc_less(o, m)) m = o;
    }
    if (lane == 0) s_d[warp] = m.d, s_id[warp] = m.id;
    __syncthreads();
    if (warp == 0) {
      ScKey w;
      w.d = s_d[lane], w.id = s_id[lane];
#pragma unroll
      for (int off = 16; off > 0; off >>= 1) {
        ScKey o;
        o.d = __shfl_xor_sync(0xffffffffu, w.d, off);
        o.id = __shfl_xor_sync(0xffffffffu, w.id, off);
        if (sc_less(o, w)) w = o;
      }
      if (lane == 0) {
        s_win = w.id;
        const bool have = w.id != INT_MAX;
        o_dist[round] = have ? __longlong_as_double((long long)w.d) : __longlong_as_double(0x7ff0000000000000ll);
        o_id[round] = have ? w.id + id_offset : -1;
        o_shift[round] = have ? shift[w.id] : 0;
      }
    }
    __syncthreads();
    if (best[0].id == s_win && s_win != INT_MAX) {
#pragma unroll
      for (int s = 0; s < kTopKMax - 1; ++s) best[s] = best[s + 1];
      best[kTopKMax - 1].d = ~0ull, best[kTopKMax - 1].id = INT_MAX;
    }
    __syncthreads();
  }
}

// Merge of gathered per-shard top-k lists on the device (what every rank runs after the all-gather).  Packed layout
// per shard: k x f64 distance | k x i32 id | k x i32 shift (16 k bytes).  One warp; round j picks the smallest
// (distance, id) strictly above round j-1's winner (ids are unique across shards), so the result is the same
// ascending (distance, id) order as ilsm_sc_merge_topk on the host.
__global__ void sc_merge_kernel(const unsigned char* __restrict__ packed, int shards, int k, unsigned char* __restrict__ out) {
  const int lane = threadIdx.x;
  const size_t rec = (size_t)16 * k;
  u64 last_d = 0;
  int last_id = -1;
  bool first = true;
  double* o_dist = reinterpret_cast<double*>(out);
  int* o_id = reinterpret_cast<int*>(out + (size_t)8 * k);
  int* o_shift = reinterpret_cast<int*>(out + (size_t)12 * k);
  for (int j = 0; j < k; ++j) {
    u64 bd = ~0ull;
    int bid = INT_MAX, bsh = 0;
    for (int e = lane; e < shards * k; e += 32) {
      const unsigned char* base = packed + (size_t)(e / k) * rec;
      const int i = e % k;
      const int id = reinterpret_cast<const int*>(base + (size_t)8 * k)[i];
      if (id < 0) continue;
      const u64 d = (u64)__double_as_longlong(reinterpret_cast<const double*>(base)[i]);
      const bool above = first || d > last_d || (d == last_d && id > last_id);
      if (above && (d < bd || (d == bd && id < bid))) bd = d, bid = id, bsh = reinterpret_cast<const int*>(base + (size_t)12 * k)[i];
    }
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) {
      const u64 od = __shfl_xor_sync(0xffffffffu, bd, off);
      const int oi = __shfl_xor_sync(0xffffffffu, bid, off);
      const int os = __shfl_xor_sync(0xffffffffu, bsh, off);
      if (od < bd || (od == bd && oi < bid)) bd = od, bid = oi, bsh = os;
    }
    const bool have = bid != INT_MAX;
    if (lane == 0) {
      o_dist[j] = have ? __longlong_as_double((long long)bd) : __longlong_as_double(0x7ff0000000000000ll);
      o_id[j] = have ? bid : -1;
      o_shift[j] = have ? bsh : 0;
    }
    if (have) last_d = bd, last_id = bid, first = false;
    else last_d = ~0ull, last_id = INT_MAX, first = false;
  }
}

int sc_merge_dev(Ctx* ctx, const void* d_packed, int shards, int k, void* d_out) {
  sc_merge_kernel<<<1, 32, 0, ctx->stream>>>(reinterpret_cast<const unsigned char*>(d_packed), shards, k,
                                             reinterpret_cast<unsigned char*>(d_out));
  count_launches(1);
  return check_launch("sc_merge");
}

// ---------------------------------------------------------------------------------------------------
// host side
// ---------------------------------------------------------------------------------------------------
int ScDb::append_dev(const float* d_desc, int n_add, bool from_host) {
  if (n_add <= 0) return ILSM_OK;
  const size_t need = (size_t)(count + n_add) * kDesc;
  if (need > db.cap) {  // grow and keep the contents
    DevBuf<float> bigger;
    int rc = bigger.reserve(need * 2);
    if (rc) return rc;
    if (count > 0)
      ILSM_CUDA(cudaMemcpyAsync(bigger.p, db.p, (size_t)count * kDesc * sizeof(float), cudaMemcpyDeviceToDevice, ctx->stream));
    ILSM_CUDA(cudaStreamSynchronize(ctx->stream));
    db.release();
    db = bigger;
  }
  ILSM_CUDA(cudaMemcpyAsync(db.p + (size_t)count * kDesc, d_desc, (size_t)n_add * kDesc * sizeof(float),
                            from_host ? cudaMemcpyHostToDevice : cudaMemcpyDeviceToDevice, ctx->stream));
  count += n_add;
  return ILSM_OK;
}

int ScDb::make_dev(const float* d_pts, int n, int stride_bytes, float* d_desc) {
  int rc;
  if ((rc = bins.reserve(kDesc))) return rc;
  cudaStream_t s = ctx->stream;
  sc_make_clear_kernel<<<(kDesc + 255) / 256, 256, 0, s>>>(bins.p);
  if (n > 0) sc_make_kernel<<<(n + 255) / 256, 256, 0, s>>>(d_pts, n, stride_bytes / 4, bins.p);
  sc_make_finish_kernel<<<(kDesc + 255) / 256, 256, 0, s>>>(bins.p, d_desc);
  count_launches(n > 0 ? 3 : 2);
  return check_launch("sc_make");
}

int ScDb::query_dev(const float* d_qdesc, int n_search, int id_offset, int k, double* d_dist, int* d_id, int* d_shift) {
  if (k < 1 || k > kTopKMax) return fail(ILSM_ERR_INVALID_ARG, "sc_query: k must be in [1,16]");
  if (n_search < 0 || n_search > count) return fail(ILSM_ERR_INVALID_ARG, "sc_query: n_search exceeds the database");
  int rc;
  if ((rc = query.reserve(1)) || (rc = dist.reserve(n_search + 1)) || (rc = shift.reserve(n_search + 1))) return rc;
  cudaStream_t s = ctx->stream;
  sc_query_prep_kernel<<<1, 64, 0, s>>>(d_qdesc, query.p);
  if (n_search > 0) {
    long long blocks = ((long long)n_search + kScWarps - 1) / kScWarps, cap = (long long)ctx->sm_count * 4;
    if (blocks > cap) blocks = cap;
    sc_score_kernel<<<(unsigned)blocks, kScWarps * 32, 0, s>>>(db.p, n_search, query.p, dist.p, shift.p);
    count_launches(1);
  }
  sc_topk_kernel<<<1, 1024, 0, s>>>(dist.p, shift.p, n_search, id_offset, k, d_dist, d_id, d_shift);
  count_launches(2);
  return check_launch("sc_query");
}

}  // namespace ilsm
