// ilsm_oracle_sc.cpp -- CPU ORACLE (TEST INFRASTRUCTURE) for ScanContext (Scancontext.cpp:25-344,
// Scancontext.h:77-96).  Descriptors are Eigen::MatrixXd 20x60 column-major in the reference; here they are stored
// row-major double[20][60] -- every value is a float widened to double (SCPointType is float), so float32 storage
// on the GPU side is lossless.  PARITY STATUS: PINNED against the reference's own code -- src/Scancontext.cpp and
// include/Scancontext.h compile unmodified against a small Eigen::MatrixXd stand-in (oracle/shims_sc) into
// oracle/_ref/libref_scancontext.so together with the vendored nanoflann: descriptor, keys, distanceBtnScanContext and
// detectLoopClosureID agree exactly (tests/test_ref_scancontext_cpu.py, tests/golden/scancontext_reference.npz).  What the
// stand-in cannot pin is the summation order inside Eigen's vectorised norm / dot / mean (left to right here and there).
#include <algorithm>
#include <cmath>
#include <cstdint>
#include <cstring>
#include <vector>

#define ORC_API extern "C" __attribute__((visibility("default")))

namespace {
constexpr int NR = 20, NS = 60;
constexpr double LIDAR_HEIGHT = 2.0, PC_MAX_RADIUS = 80.0, SEARCH_RATIO = 0.1;

// Scancontext.cpp:25-42 (bitwise & on bools: same truth table for finite inputs)
// atan of a float argument: evaluated in double and rounded to float (= correctly rounded atanf; the reference's
// unqualified atan() resolves to the float or the double overload depending on which headers ROS pulls in).
inline float atan_f(float v) { return (float)std::atan((double)v); }
float xy2theta(const float& x, const float& y) {
  if (x >= 0 && y >= 0) return (float)((180 / M_PI) * atan_f(y / x));
  if (x < 0 && y >= 0) return (float)(180 - ((180 / M_PI) * atan_f(y / (-x))));
  if (x < 0 && y < 0) return (float)(180 + ((180 / M_PI) * atan_f(y / x)));
  return (float)(360 - ((180 / M_PI) * atan_f((-y) / x)));
}

inline double at(const double* d, int r, int c) { return d[r * NS + c]; }

// distDirectSC(sc1, circshift(sc2, shift)) (Scancontext.cpp:44-68,79-101): shifted column j = original column (j - shift) mod 60
double dist_direct(const double* a, const double* b, int shift) {
  int num_eff = 0;
  double sum = 0;
  for (int j = 0; j < NS; ++j) {
    const int jb = ((j - shift) % NS + NS) % NS;
    double na = 0, nb = 0, dot = 0;
    for (int r = 0; r < NR; ++r) {
      na += at(a, r, j) * at(a, r, j);
      nb += at(b, r, jb) * at(b, r, jb);
      dot += at(a, r, j) * at(b, r, jb);
    }
    na = std::sqrt(na), nb = std::sqrt(nb);
    if (na == 0 || nb == 0) continue;
    sum = sum + dot / (na * nb);
    num_eff = num_eff + 1;
  }
  if (num_eff == 0) return INFINITY;  // the reference divides 0/0 here; defined as "no match"
  return 1.0 - sum / num_eff;
}

void sector_key(const double* d, double* key) {
  for (int c = 0; c < NS; ++c) {
    double s = 0;
    for (int r = 0; r < NR; ++r) s += at(d, r, c);
    key[c] = s / NR;
  }
}

int fast_align(const double* k1, const double* k2) {
  int arg = 0;
  double best = 10000000;
  for (int s = 0; s < NS; ++s) {
    double ss = 0;
    for (int c = 0; c < NS; ++c) {
      const double diff = k1[c] - k2[((c - s) % NS + NS) % NS];
      ss += diff * diff;
    }
    const double nrm = std::sqrt(ss);
    if (nrm < best) best = nrm, arg = s;
  }
  return arg;
}

// distanceBtnScanContext (Scancontext.cpp:126-157)
void sc_distance(const double* q, const double* c, double* dist, int* shift) {
  double kq[NS], kc[NS];
  sector_key(q, kq);
  sector_key(c, kc);
  const int a = fast_align(kq, kc);
  const int radius = (int)std::round(0.5 * SEARCH_RATIO * NS);
  std::vector<int> space{a};
  for (int ii = 1; ii < radius + 1; ++ii) {
    space.push_back((a + ii + NS) % NS);
    space.push_back((a - ii + NS) % NS);
  }
  std::sort(space.begin(), space.end());
  int arg = 0;
  double best = 10000000;
  for (int s : space) {
    const double d = dist_direct(q, c, s);
    if (d < best) best = d, arg = s;
  }
  *dist = best, *shift = arg;
}
}  // namespace

// makeScancontext (Scancontext.cpp:160-204): desc row-major double[20*60]
ORC_API void orc_sc_make(const float* pts, int n, int stride_bytes, double* desc) {
  const int sf = stride_bytes / 4;
  const double NO_POINT = -1000;
  for (int i = 0; i < NR * NS; ++i) desc[i] = NO_POINT;
  for (int i = 0; i < n; ++i) {
    const float* p = pts + (size_t)i * sf;
    const float x = p[0], y = p[1];
    const float z = (float)(p[2] + LIDAR_HEIGHT);
    const float azim_range = std::sqrt(x * x + y * y);
    const float azim_angle = xy2theta(x, y);
    if (azim_range > PC_MAX_RADIUS) continue;
    const int ring = std::max(std::min(NR, int(std::ceil((azim_range / PC_MAX_RADIUS) * NR))), 1);
    const int sector = std::max(std::min(NS, int(std::ceil((azim_angle / 360.0) * NS))), 1);
    double& cell = desc[(ring - 1) * NS + (sector - 1)];
    if (cell < z) cell = z;
  }
  for (int i = 0; i < NR * NS; ++i)
    if (desc[i] == NO_POINT) desc[i] = 0;
}

// makeRingkeyFromScancontext / makeSectorkeyFromScancontext (Scancontext.cpp:206-235)
ORC_API void orc_sc_keys(const double* desc, double* ring_key20, double* sector_key60) {
  for (int r = 0; r < NR; ++r) {
    double s = 0;
    for (int c = 0; c < NS; ++c) s += at(desc, r, c);
    ring_key20[r] = s / NS;
  }
  sector_key(desc, sector_key60);
}

ORC_API void orc_sc_distance(const double* q, const double* c, double* dist, int32_t* shift) {
  int s;
  sc_distance(q, c, dist, &s);
  *shift = s;
}

// brute-force scoring of a query against db[0..n) and top-k by (distance, id)
ORC_API void orc_sc_topk(const double* db, int n, const double* q, int k, double* dist, int32_t* id, int32_t* shift) {
  struct R {
    double d;
    int id, sh;
  };
  std::vector<R> all(n);
  for (int i = 0; i < n; ++i) {
    sc_distance(q, db + (size_t)i * NR * NS, &all[i].d, &all[i].sh);
    all[i].id = i;
  }
  std::sort(all.begin(), all.end(), [](const R& a, const R& b) { return a.d < b.d || (a.d == b.d && a.id < b.id); });
  for (int j = 0; j < k; ++j) {
    if (j < n) {
      dist[j] = all[j].d, id[j] = all[j].id, shift[j] = all[j].sh;
    } else {
      dist[j] = INFINITY, id[j] = -1, shift[j] = 0;
    }
  }
}
