// scancontext_tc.cu -- K5 at scale: batched ScanContext scoring as the contraction it is (BASELINE configs[4]).
//
// The exact scorer (scancontext.cu: fp64 like Eigen, one warp per candidate) is bound by fp64 latency and by
// shared-memory loads: every staged value feeds ONE multiply-add.  Scoring a 100k-keyframe database that way costs
// 0.46 ms per query.  Almost all of that work decides nothing: only the k best candidates are reported.  So:
//
//   1. sc_prefilter_kernel -- APPROXIMATE distance of every (query, candidate) pair on the tensor cores, 8 queries per
//      block, the candidate staged once for all of them:
//        * sector-key alignment (fastAlignUsingVkey, Scancontext.cpp:104-124): argmin_s |qk - shift(ck, s)| = argmax_s of the
//          circular cross-correlation, a (64 shifts x 64 columns) circulant of the candidate key times the (64 x 8) matrix
//          of query keys: mma.sync.m16n8k16 f16 with a hi/lo split of both operands (three products: error ~2^-22) and
//          fp32 accumulation.  The two best correlations are kept: when they are closer than a rigorous bound on the
//          arithmetic error the pair is FLAGGED (its alignment is not certain) and always rescored exactly;
//        * the 7-shift column-cosine block (distDirectSC over the shifts around the aligned one, :79-101, 126-157): with
//          column-normalised descriptors the cosines are a BAND of the 60 x 60 Gram matrix Qhat^T Chat (inner dimension
//          20 rings): per 16 query columns a 16 x 24 tile product against the candidate columns starting at
//          16 i - shift - 3, six m16n8k16 per tile row, B operands by ldmatrix from a conflict-free (80-byte) column
//          layout; the seven diagonals are summed straight out of the accumulator registers.  The effective-column
//          counts are popcounts of rotated 60-bit validity masks.
//      Error of an unflagged distance: <= kPfEps (f16 rounding of the normalised entries, 2^-11 each).
//   2. selection -- per query the k-th smallest approximate distance T; every candidate with approx <= T + 2 kPfEps, and
//      every flagged one, goes to the query's list.  The list provably contains the exact top-k (any member of the exact
//      top-k has exact <= T + eps, hence approx <= T + 2 eps).
//   3. the exact fp64 kernel in list mode over those few candidates, then the exact top-k by (distance, id).
// Results are identical to scoring every entry exactly (tests compare the two), at a fraction of the fp64 work.
#include <cuda_fp16.h>

#include "ilsm_host.hpp"

namespace ilsm {

constexpr int kNR = 20, kNS = 60, kDesc = kNR * kNS;
constexpr int kPfQ = 8;            // queries per block, one warp each
constexpr int kPfThreads = 32 * kPfQ;
constexpr int kChStride = 40;      // halves per staged candidate column: 80 B rows -> ldmatrix / LDS.32 conflict-free
constexpr int kChCols = 84;        // 60 columns + the first 24 again: a 24-column window never wraps
constexpr float kPfEps = 1.5e-3f;  // bound of |approximate - exact| distance for an unflagged pair

struct PfQuery {              // prepared once per query
  __half qhat[64][32];        // [column][ring] column-normalised descriptor, zero padded (columns 60..63, rings 20..31)
  __half kh[64], kl[64];      // sector key, f16 hi / lo split (columns >= 60 zero)
  float knorm;                // |sector key|
  unsigned long long mask;    // bit c: column c has a non-zero norm
};

__device__ __forceinline__ void mma_f16(float (&d)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
  asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.f16.f16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
               : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3])
               : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}
__device__ __forceinline__ void ldmatrix_x4(uint32_t (&r)[4], const void* smem_row) {
  const uint32_t addr = (uint32_t)__cvta_generic_to_shared(smem_row);
  asm volatile("ldmatrix.sync.aligned.m8n8.x4.shared.b16 {%0,%1,%2,%3}, [%4];" : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]) : "r"(addr));
}
__device__ __forceinline__ uint32_t pack_h2(__half lo, __half hi) {
  return (uint32_t)__half_as_ushort(lo) | ((uint32_t)__half_as_ushort(hi) << 16);
}

// ---------------------------------------------------------------------------------------------------
// query preparation: one block of 64 threads per query
// ---------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(64) sc_pf_prep_kernel(const float* __restrict__ qdesc, PfQuery* __restrict__ out) {
  pdl_entry();
  const float* d = qdesc + (size_t)blockIdx.x * kDesc;
  PfQuery& q = out[blockIdx.x];
  const int c = threadIdx.x;
  float key = 0.f, inv = 0.f;
  bool valid = false;
  if (c < kNS) {
    float s = 0.f, ss = 0.f;
    for (int r = 0; r < kNR; ++r) {
      const float v = d[r * kNS + c];
      s += v;
      ss = fmaf(v, v, ss);
    }
    key = s / kNR;
    valid = ss > 0.f;
    inv = valid ? rsqrtf(ss) : 0.f;
  }
  for (int r = 0; r < 32; ++r) q.qhat[c][r] = __float2half(c < kNS && r < kNR ? d[r * kNS + c] * inv : 0.f);
  const __half h = __float2half(key);
  q.kh[c] = h;
  q.kl[c] = __float2half(key - __half2float(h));
  __shared__ float s_k2[2];
  __shared__ unsigned s_m[2];
  float k2 = key * key;
#pragma unroll
  for (int off = 16; off > 0; off >>= 1) k2 += __shfl_xor_sync(0xffffffffu, k2, off);
  const unsigned m = __ballot_sync(0xffffffffu, valid);
  if ((threadIdx.x & 31) == 0) s_k2[threadIdx.x >> 5] = k2, s_m[threadIdx.x >> 5] = m;
  __syncthreads();
  if (threadIdx.x == 0) {
    q.knorm = sqrtf(s_k2[0] + s_k2[1]);
    q.mask = (unsigned long long)s_m[0] | ((unsigned long long)s_m[1] << 32);
  }
}

__device__ __forceinline__ void mma_f16_zero(float (&d)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {  // D = A B
  asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.f16.f16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%10,%10,%10,%10};"
               : "=f"(d[0]), "=f"(d[1]), "=f"(d[2]), "=f"(d[3])
               : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1), "f"(0.f));
}
__device__ __forceinline__ uint32_t redux_max_u32(uint32_t v) {
  uint32_t r;
  asm volatile("redux.sync.max.u32 %0, %1, 0xffffffff;" : "=r"(r) : "r"(v));
  return r;
}
__device__ __forceinline__ void cp_async16(void* smem_dst, const void* gsrc) {
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"((uint32_t)__cvta_generic_to_shared(smem_dst)), "l"(gsrc) : "memory");
}
__device__ __forceinline__ void cp_async_commit_wait() {
  asm volatile("cp.async.commit_group;\n\tcp.async.wait_group 0;" ::: "memory");
}
__device__ __forceinline__ void named_barrier(int id, int threads) { asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(threads) : "memory"); }

// ---------------------------------------------------------------------------------------------------
// the prefilter
// ---------------------------------------------------------------------------------------------------
struct PfSmem {
  __align__(16) float craw[2][kDesc];           // raw candidate descriptors, double buffered (filled by cp.async)
  __align__(16) __half ch[kChCols][kChStride];  // column-normalised candidate, [column][ring] (columns 60.. repeat 0..); rings 20..31 stay zero
  uint32_t pkh[128], pkl[128];                  // candidate sector key hi / lo as adjacent pairs (x, x + 1), x = (column - shift) + 64, periodic
  float corr[2][64][kPfQ];                      // alignment correlations [column half][shift][query] (the two halves are added by the reader)
  float ck2_part[8];
  unsigned cm_part[8];
  float ck2;
  unsigned long long cmask;
};

// D[q * n + c] = approximate distance of query q and candidate c; -1 when the pair is flagged for exact rescoring; <= -2 when
// its alignment is one of two (the value is -2 - the smaller of the two distances, a lower bound of the exact one);
// Sh (optional, debugging / tests) = the aligned shift the prefilter used.
__global__ void __launch_bounds__(kPfThreads, 3)
    sc_prefilter_kernel(const float* __restrict__ db, int n, const PfQuery* __restrict__ Q, int B, float* __restrict__ D,
                        unsigned char* __restrict__ Sh) {
  pdl_entry();
  extern __shared__ __align__(16) unsigned char pf_raw[];
  PfSmem& sm = *reinterpret_cast<PfSmem*>(pf_raw);
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int g = lane >> 2, t = lane & 3;
  {  // blockIdx.y = which group of 8 queries of the batch this block serves
    const int q0 = (int)blockIdx.y * kPfQ;
    Q += q0, D += (size_t)q0 * n, B = B - q0 < kPfQ ? B - q0 : kPfQ;
    if (Sh) Sh += (size_t)q0 * n;
  }
  const bool has_q = warp < B;
  const PfQuery& myq = Q[has_q ? warp : 0];

  // ---- per-warp constants: A fragments of the normalised query (cosine band), 4 tile rows x 2 k-steps
  uint32_t qa[4][2][4];
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int ks = 0; ks < 2; ++ks) {
      const __half* r0 = &myq.qhat[16 * i + g][16 * ks + 2 * t];
      const __half* r1 = &myq.qhat[16 * i + g + 8][16 * ks + 2 * t];
      qa[i][ks][0] = *reinterpret_cast<const uint32_t*>(r0);
      qa[i][ks][1] = *reinterpret_cast<const uint32_t*>(r1);
      qa[i][ks][2] = *reinterpret_cast<const uint32_t*>(r0 + 8);
      qa[i][ks][3] = *reinterpret_cast<const uint32_t*>(r1 + 8);
    }
  const unsigned long long qmask = myq.mask;
  const float qknorm = myq.knorm;
  // ---- alignment: warp w owns shifts 16 (w & 3) .. + 15 and columns 32 (w >> 2) .. + 31 (two of the four k-steps);
  // B fragments = the sector keys of the block's queries, query index = g
  const int am = warp & 3, ak = warp >> 2;
  uint32_t kb_h[2][2], kb_l[2][2];
  {
    const bool qv = g < B;
    const PfQuery& kq = Q[qv ? g : 0];
#pragma unroll
    for (int kk = 0; kk < 2; ++kk) {
      const int c0 = 16 * (2 * ak + kk) + 2 * t;
      kb_h[kk][0] = qv ? *reinterpret_cast<const uint32_t*>(&kq.kh[c0]) : 0u;
      kb_h[kk][1] = qv ? *reinterpret_cast<const uint32_t*>(&kq.kh[c0 + 8]) : 0u;
      kb_l[kk][0] = qv ? *reinterpret_cast<const uint32_t*>(&kq.kl[c0]) : 0u;
      kb_l[kk][1] = qv ? *reinterpret_cast<const uint32_t*>(&kq.kl[c0 + 8]) : 0u;
    }
  }
  // epilogue geometry of the cosine band (see the header): accumulator (row m, column o) of a 16 x 24 tile belongs to
  // shift index tt = m + 6 - o of the seven; per lane that is tt0 = u & 7 for its even columns and tt0 - 1 for the odd
  // ones, from the n-tile j = h (u <= 7) or j = h + 1 (u >= 8), h = which row half
  const int u = g - 2 * t + 6;
  const bool upper = u >= 8;
  const int tt0 = u & 7;

  // the padding rings of the staged candidate are written once
  for (int i = tid; i < kChCols * (kChStride - kNR); i += kPfThreads) sm.ch[i / (kChStride - kNR)][kNR + i % (kChStride - kNR)] = __float2half(0.f);
  // first candidate of this block (4800 B = 300 x 16 B)
  if ((int)blockIdx.x < n)
    for (int i = tid; i < kDesc / 4; i += kPfThreads) cp_async16(&sm.craw[0][4 * i], db + (size_t)blockIdx.x * kDesc + 4 * i);
  int cur = 0;
#pragma unroll 1
  for (int cand = blockIdx.x; cand < n; cand += gridDim.x) {
    cp_async_commit_wait();
    __syncthreads();  // craw[cur] complete; everybody is done with the previous candidate's staged data
    // ---- the next candidate streams into the other buffer while this one is processed
    const int nxt = cand + gridDim.x;
    if (nxt < n)
      for (int i = tid; i < kDesc / 4; i += kPfThreads) cp_async16(&sm.craw[cur ^ 1][4 * i], db + (size_t)nxt * kDesc + 4 * i);
    // ---- stage (all 8 warps: 4 threads per column, rings 0-5 / 6-11 / 12-15 / 16-19): column sums, normalise, packed
    // f16 stores; sector key hi / lo; validity mask; |key|^2
    {
      const int col = tid >> 2, q4 = tid & 3;  // 240 threads work (col < 60)
      const int r0 = q4 < 2 ? 6 * q4 : 4 + 4 * q4, cnt = q4 < 2 ? 6 : 4;
      float v[6], s = 0.f, ss = 0.f;
#pragma unroll
      for (int k = 0; k < 6; ++k) {
        v[k] = (col < kNS && k < cnt) ? sm.craw[cur][(r0 + k) * kNS + col] : 0.f;
        s += v[k];
        ss = fmaf(v[k], v[k], ss);
      }
      s += __shfl_xor_sync(0xffffffffu, s, 1), ss += __shfl_xor_sync(0xffffffffu, ss, 1);
      s += __shfl_xor_sync(0xffffffffu, s, 2), ss += __shfl_xor_sync(0xffffffffu, ss, 2);
      const bool valid = ss > 0.f;
      const float inv = valid ? rsqrtf(ss) : 0.f;
      const float key = s / kNR;
      if (col < kNS) {
        uint32_t* dst = reinterpret_cast<uint32_t*>(&sm.ch[col][r0]);
        uint32_t* dup = reinterpret_cast<uint32_t*>(&sm.ch[col < kChCols - kNS ? col + kNS : col][r0]);
#pragma unroll
        for (int k = 0; k < 3; ++k) {
          if (2 * k < cnt) {
            const __half2 h2 = __floats2half2_rn(v[2 * k] * inv, v[2 * k + 1] * inv);
            dst[k] = *reinterpret_cast<const uint32_t*>(&h2);
            dup[k] = *reinterpret_cast<const uint32_t*>(&h2);
          }
        }
        if (q4 == 0) {
          const __half h = __float2half(key), l = __float2half(key - __half2float(h));
          // index x = (column - shift) + 64 holds column (x + 56) mod 60; the keys are kept as adjacent PAIRS
          // pk[x] = (key[x], key[x + 1]), so that an A-fragment register of the circulant is one 32-bit load: key[x] is
          // the low half of pk[x] and the high half of pk[x - 1]
          __half* ph = reinterpret_cast<__half*>(sm.pkh);
          __half* pl = reinterpret_cast<__half*>(sm.pkl);
          auto put = [&](int x) {
            if (x < 128) ph[2 * x] = h, pl[2 * x] = l;
            if (x > 0) ph[2 * x - 1] = h, pl[2 * x - 1] = l;
          };
          put(col + 4), put(col + 64);
          if (col < 4) put(col + 124);
          if (col == 4) put(128);  // only the high half of pk[127]
          if (col >= 56) put(col - 56);
        }
      }
      const unsigned vb = __ballot_sync(0xffffffffu, valid && q4 == 0 && col < kNS);
      float k2 = (q4 == 0 && col < kNS) ? key * key : 0.f;
#pragma unroll
      for (int off = 16; off > 0; off >>= 1) k2 += __shfl_xor_sync(0xffffffffu, k2, off);
      if (lane == 0) {
        unsigned m8 = 0;  // lanes 0, 4, .., 28 of warp w carry columns 8 w .. 8 w + 7
#pragma unroll
        for (int k = 0; k < 8; ++k) m8 |= ((vb >> (4 * k)) & 1u) << k;
        sm.cm_part[warp] = m8;
        sm.ck2_part[warp] = k2;
      }
      __syncthreads();
      if (tid == 255) {  // (read after the next barrier, by the band phase)
        sm.ck2 = ((sm.ck2_part[0] + sm.ck2_part[1]) + (sm.ck2_part[2] + sm.ck2_part[3])) +
                 ((sm.ck2_part[4] + sm.ck2_part[5]) + (sm.ck2_part[6] + sm.ck2_part[7]));
        unsigned long long m = 0ull;
#pragma unroll
        for (int w = 0; w < 8; ++w) m |= (unsigned long long)sm.cm_part[w] << (8 * w);
        sm.cmask = m & ((1ull << kNS) - 1ull);
      }
      // ---- alignment: the correlations of shifts 16 am .. + 15 with all queries over columns 32 ak .. + 31
      float acc[4];
#pragma unroll
      for (int kk = 0; kk < 2; ++kk) {
        const int ks = 2 * ak + kk;
        // A[m][k] = ck[(column - shift) mod 60], shift = 16 am + m, column = 16 ks + k  ->  index x = column - shift + 64
        const int x0 = 16 * ks + 2 * t - (16 * am + g) + 64;
        uint32_t ah[4], al[4];
        ah[0] = sm.pkh[x0], al[0] = sm.pkl[x0];
        ah[1] = sm.pkh[x0 - 8], al[1] = sm.pkl[x0 - 8];  // shift + 8
        ah[2] = sm.pkh[x0 + 8], al[2] = sm.pkl[x0 + 8];  // column + 8
        ah[3] = ah[0], al[3] = al[0];                    // both
        if (kk == 0) mma_f16_zero(acc, ah, kb_h[kk][0], kb_h[kk][1]);
        else mma_f16(acc, ah, kb_h[kk][0], kb_h[kk][1]);
        mma_f16(acc, ah, kb_l[kk][0], kb_l[kk][1]);
        mma_f16(acc, al, kb_h[kk][0], kb_h[kk][1]);
      }
      // acc[0]: (shift 16 am + g, query 2 t), acc[1]: (same shift, query 2 t + 1), acc[2] / acc[3]: shift + 8
      const int s0 = 16 * am + g;
      *reinterpret_cast<float2*>(&sm.corr[ak][s0][2 * t]) = make_float2(acc[0], acc[1]);
      *reinterpret_cast<float2*>(&sm.corr[ak][s0 + 8][2 * t]) = make_float2(acc[2], acc[3]);
    }
    __syncthreads();
    // ---- every warp: its own query against the staged candidate
    if (has_q) {
      // best and second best correlation over the 60 shifts: order-preserving integer keys (correlations of non-negative
      // keys are >= 0 up to rounding) with the shift in the low 6 bits, two REDUX each.  Truncating 6 mantissa bits
      // costs 2^-17 relative, accounted for in the bound below.
      const float c0v = sm.corr[0][lane][warp] + sm.corr[1][lane][warp];
      const float c1v = lane + 32 < kNS ? sm.corr[0][lane + 32][warp] + sm.corr[1][lane + 32][warp] : 0.f;
      uint32_t k0 = (__float_as_uint(fmaxf(c0v, 0.f)) & ~63u) | (uint32_t)(63 - lane);          // lower shift wins ties
      uint32_t k1 = lane + 32 < kNS ? (__float_as_uint(fmaxf(c1v, 0.f)) & ~63u) | (uint32_t)(31 - lane) : 0u;
      const uint32_t best = redux_max_u32(k0 > k1 ? k0 : k1);
      if (k0 == best) k0 = 0u;
      if (k1 == best) k1 = 0u;
      const uint32_t second = redux_max_u32(k0 > k1 ? k0 : k1);
      if (k0 == second) k0 = 0u;
      if (k1 == second) k1 = 0u;
      const uint32_t third = redux_max_u32(k0 > k1 ? k0 : k1);
      const int a1 = 63 - (int)(best & 63u), a2 = 63 - (int)(second & 63u);
      const float b1 = __uint_as_float(best & ~63u), b2 = __uint_as_float(second & ~63u), b3 = __uint_as_float(third & ~63u);
      const unsigned long long cmask = sm.cmask;
      // |correlation error| <= 3e-5 |qk| |ck| (hi/lo split: dropped lo x lo term 2^-22; fp32 accumulation of 192 products
      // 1.1e-5; key truncation 7.6e-6); the alignment is certain when the best beats the runner-up by more than twice that.
      // When only the runner-up is that close, the true alignment is one of the two: the band is evaluated for both and
      // the pair carries the smaller distance as a LOWER bound (dual); three or more contenders: flagged outright.
      const float err = 3.0e-5f * qknorm * sqrtf(sm.ck2) + 1e-30f;
      const bool unsure = !(b1 - b2 > 2.f * err);
      const bool hard = a1 >= kNS || (unsure && (a2 >= kNS || !(b1 - b3 > 2.f * err)));
      const bool dual = unsure && !hard;
      // the band for one alignment (inlined twice: the second copy only runs for the rare dual pairs)
      auto band = [&](const int a) -> float {
      float x0 = 0.f, x1 = 0.f;
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        int cb = 16 * i - a - 3;  // first candidate column of the tile row's window (24 columns, no wrap in the periodic copy)
        cb += cb < 0 ? kNS : 0;
        cb += cb < 0 ? kNS : 0;   // 16 i - a - 3 >= -62
        const __half* rowp = &sm.ch[cb + (lane & 7)][8 * (lane >> 3)];
        float d[3][4];
#pragma unroll
        for (int j = 0; j < 3; ++j) {
          uint32_t b[4];
          ldmatrix_x4(b, rowp + 8 * j * kChStride);
          mma_f16_zero(d[j], qa[i][0], b[0], b[1]);
          mma_f16(d[j], qa[i][1], b[2], b[3]);
        }
        x0 += upper ? (d[1][0] + d[2][2]) : (d[0][0] + d[1][2]);
        x1 += upper ? (d[1][1] + d[2][3]) : (d[0][1] + d[1][3]);
      }
      // sum over the four lanes that share tt0 (one per t): lane (g', t') with g' = (g - 2 t + 2 t') mod 8
      float X0 = 0.f, X1 = 0.f;
#pragma unroll
      for (int tp = 0; tp < 4; ++tp) {
        const int src = 4 * ((g - 2 * t + 2 * tp) & 7) + tp;
        X0 += __shfl_sync(0xffffffffu, x0, src);
        X1 += __shfl_sync(0xffffffffu, x1, src);
      }
      // shift index tt0 collects its even-column sum and the odd-column sum of the lanes with tt0 + 1
      const float X1n = __shfl_sync(0xffffffffu, X1, 4 * ((g + 1) & 7) + t);
      float dist = 3.0e38f;
      if (tt0 <= 6) {
        int s = a - 3 + tt0;
        s += s < 0 ? kNS : 0;
        s -= s >= kNS ? kNS : 0;
        // column j of the query meets column (j - s) mod 60 of the candidate: candidate mask rotated left by s
        const unsigned long long rot = ((cmask << s) | (cmask >> (kNS - s))) & ((1ull << kNS) - 1ull);
        const int eff = __popcll(qmask & (s ? rot : cmask));
        if (eff > 0) dist = 1.f - (X0 + X1n) / (float)eff;
      }
#pragma unroll
      for (int off = 16; off > 0; off >>= 1) dist = fminf(dist, __shfl_xor_sync(0xffffffffu, dist, off));
      return dist;
      };
      float dbest = band(a1 < kNS ? a1 : 0);
      if (dual) dbest = fminf(dbest, band(a2));
      if (lane == 0) {
        // >= 0: distance (error <= kPfEps);  -1: flagged;  <= -2: dual, -2 - (lower bound of the distance)
        D[(size_t)warp * n + cand] = hard ? -1.f : (dual ? -2.f - dbest : dbest);
        if (Sh) Sh[(size_t)warp * n + cand] = (unsigned char)(a1 < kNS ? a1 : 0);
      }
    }
    cur ^= 1;
  }
}

// ---------------------------------------------------------------------------------------------------
// selection: k-th smallest approximate distance per query, then the rescoring list
// ---------------------------------------------------------------------------------------------------
__device__ __forceinline__ void block_k_smallest_u64(u64 mine, int k, u64* __restrict__ out) {
  __shared__ u64 s_w[32];
  __shared__ u64 s_win;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nw = blockDim.x >> 5;
  for (int round = 0; round < k; ++round) {
    u64 m = mine;
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) {
      const u64 o = __shfl_xor_sync(0xffffffffu, m, off);
      m = o < m ? o : m;
    }
    if (lane == 0) s_w[warp] = m;
    __syncthreads();
    if (warp == 0) {
      u64 w = lane < nw ? s_w[lane] : ~0ull;
#pragma unroll
      for (int off = 16; off > 0; off >>= 1) {
        const u64 o = __shfl_xor_sync(0xffffffffu, w, off);
        w = o < w ? o : w;
      }
      if (lane == 0) s_win = w, out[round] = w;
    }
    __syncthreads();
    if (mine == s_win) mine = ~0ull;
    __syncthreads();
  }
}

// grid (chunks, B): the k smallest unflagged approximate distances of each 8192-candidate chunk
constexpr int kSelChunk = 8192;
__global__ void __launch_bounds__(1024) sc_pf_select1_kernel(const float* __restrict__ D, int n, int k, u64* __restrict__ part) {
  pdl_entry();
  const float* d = D + (size_t)blockIdx.y * n;
  const int base = blockIdx.x * kSelChunk;
  // every thread keeps the smallest of its 8 strided entries; k rounds over the block then refill from the rest would be
  // exact, but a chunk's k smallest may sit in one thread: so every thread offers ALL its entries, 8 passes of k rounds
  u64* out = part + ((size_t)blockIdx.y * gridDim.x + blockIdx.x) * (size_t)(8 * k);
  for (int e = 0; e < 8; ++e) {
    const int i = base + e * 1024 + threadIdx.x;
    u64 mine = ~0ull;
    if (i < n) {
      const float v = d[i];
      if (v >= 0.f) mine = ((u64)__float_as_uint(v) << 32) | (uint32_t)i;
    }
    block_k_smallest_u64(mine, k, out + e * k);
  }
}
// one block per query: threshold = k-th smallest of the chunk lists + 2 eps  (+inf when fewer than k unflagged exist)
__global__ void __launch_bounds__(1024) sc_pf_select2_kernel(const u64* __restrict__ part, int m, int k, float* __restrict__ thr, int* __restrict__ list_n) {
  pdl_entry();
  __shared__ u64 s_out[16];
  __shared__ u64 s_last;
  const u64* p = part + (size_t)blockIdx.x * m;
  if (threadIdx.x == 0) s_last = 0;
  __syncthreads();
  for (int round = 0; round < k; ++round) {
    const u64 last = s_last;
    u64 mine = ~0ull;
    for (int e = threadIdx.x; e < m; e += blockDim.x) {
      const u64 v = p[e];
      if ((round == 0 || v > last) && v < mine) mine = v;
    }
    __syncthreads();
    block_k_smallest_u64(mine, 1, s_out + round);
    if (threadIdx.x == 0) s_last = s_out[round];
    __syncthreads();
  }
  if (threadIdx.x == 0) {
    const u64 kth = s_out[k - 1];
    thr[blockIdx.x] = kth == ~0ull ? __int_as_float(0x7f800000) : __uint_as_float((uint32_t)(kth >> 32)) + 2.f * kPfEps;
    list_n[blockIdx.x] = 0;
  }
}
// grid (blocks, B): candidates at or below the threshold, and the flagged ones, into the query's list (order arbitrary:
// the exact top-k afterwards is by (distance, id))
__global__ void __launch_bounds__(256) sc_pf_compact_kernel(const float* __restrict__ D, int n, const float* __restrict__ thr, u64* __restrict__ list,
                                                            int* __restrict__ list_n) {  // list capacity = n per query: cannot overflow
  pdl_entry();
  const int q = blockIdx.y;
  const float th = thr[q];
  const float* d = D + (size_t)q * n;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
    const float v = d[i];
    // flagged pairs always; dual pairs by their lower bound; the rest by their distance
    if (v == -1.f || (v <= -2.f ? -2.f - v <= th : (v >= 0.f && v <= th))) {
      const int pos = atomicAdd(&list_n[q], 1);
      list[(size_t)q * n + pos] = (u64)(uint32_t)i;
    }
  }
}

// ---------------------------------------------------------------------------------------------------
// host side
// ---------------------------------------------------------------------------------------------------
int sc_score_lists_dev(ScDb* db, const float* d_qdesc, int B, int k, const u64* d_lists, const int* d_list_n, int list_cap, int id_offset,
                       int n_search, unsigned char* d_packed);  // scancontext.cu

int ScDb::query_batch_tc_dev(const float* d_qdesc, int B, int n_search, int id_offset, int k, unsigned char* d_packed, unsigned char* d_shift_dbg) {
  if (k < 1 || k > 16) return fail(ILSM_ERR_INVALID_ARG, "sc_query: k must be in [1,16]");
  if (n_search < 0 || n_search > count) return fail(ILSM_ERR_INVALID_ARG, "sc_query: n_search exceeds the database");
  cudaStream_t s = ctx->stream;
  int rc;
  // the whole batch goes through every step in ONE launch each (the prefilter serves groups of 8 queries through
  // blockIdx.y): at small shards (a database split over 8 GPUs) the per-launch costs are what a query batch pays for
  constexpr int kBatchMax = 64;
  for (int b0 = 0; b0 < B; b0 += kBatchMax) {
    const int nb = B - b0 < kBatchMax ? B - b0 : kBatchMax;
    const int groups = (nb + kPfQ - 1) / kPfQ;
    const int chunks = n_search > 0 ? (n_search + kSelChunk - 1) / kSelChunk : 1;
    const size_t part_per_q = (size_t)chunks * 8 * k;
    if ((rc = pf_query.reserve((size_t)groups * kPfQ * sizeof(PfQuery))) || (rc = pf_dist.reserve((size_t)nb * (n_search + 1))) ||
        (rc = pf_part.reserve(nb * part_per_q + 16)) || (rc = pf_thr.reserve(nb)) || (rc = pf_list_n.reserve(nb)) ||
        (rc = pf_list.reserve((size_t)nb * (n_search + 1))))
      return rc;
    ILSM_CUDA(launch_pdl(sc_pf_prep_kernel, dim3(nb), dim3(64), 0, s, d_qdesc + (size_t)b0 * kDesc, reinterpret_cast<PfQuery*>(pf_query.p)));
    int launches = 1;
    if (n_search > 0) {
      ILSM_CUDA(cudaFuncSetAttribute(sc_prefilter_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(PfSmem)));
      int blocks = ctx->sm_count * 3;
      if (blocks > n_search) blocks = n_search;
      ILSM_CUDA(launch_pdl(sc_prefilter_kernel, dim3(blocks, groups), dim3(kPfThreads), sizeof(PfSmem), s, (const float*)db.p, n_search,
                           (const PfQuery*)pf_query.p, nb, pf_dist.p, d_shift_dbg));
      ++launches;
    }
    ILSM_CUDA(launch_pdl(sc_pf_select1_kernel, dim3(chunks, nb), dim3(1024), 0, s, (const float*)pf_dist.p, n_search, k, pf_part.p));
    ILSM_CUDA(launch_pdl(sc_pf_select2_kernel, dim3(nb), dim3(1024), 0, s, (const u64*)pf_part.p, (int)part_per_q, k, pf_thr.p, pf_list_n.p));
    if (n_search > 0) {
      ILSM_CUDA(launch_pdl(sc_pf_compact_kernel, dim3(ctx->sm_count, nb), dim3(256), 0, s, (const float*)pf_dist.p, n_search, (const float*)pf_thr.p,
                           pf_list.p, pf_list_n.p));
      ++launches;
    }
    count_launches(launches + 2);
    if ((rc = sc_score_lists_dev(this, d_qdesc + (size_t)b0 * kDesc, nb, k, pf_list.p, pf_list_n.p, n_search, id_offset, n_search,
                                 d_packed + (size_t)b0 * 16 * k)))
      return rc;
  }
  return check_launch("sc_query_tc");
}

}  // namespace ilsm
