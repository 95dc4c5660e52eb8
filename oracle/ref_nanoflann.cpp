// ref_nanoflann.cpp -- thin C wrapper around the REFERENCE'S OWN vendored nanoflann 1.3.2
// (/root/reference/include/nanoflann.hpp, KDTreeVectorOfVectorsAdaptor.h), the only reference source that
// compiles in this environment.  TEST INFRASTRUCTURE: it pins the oracle's k-NN and serves as the k-d tree of
// the CPU baseline.  The headers are included from where they lie (-I/root/reference/include, see Makefile);
// no reference source is copied into this repository.  Output: oracle/_ref/libref_nanoflann.so (git-ignored).
#define NANOFLANN_FIRST_MATCH 1  // include/nanoflann.hpp:177-184: equal distance -> lowest index first
#include <nanoflann.hpp>
#include <KDTreeVectorOfVectorsAdaptor.h>

#include <cstdint>
#include <memory>
#include <vector>

namespace {
struct StridedCloud {
  const float* base;
  size_t n;
  int stride_f;
  inline size_t kdtree_get_point_count() const { return n; }
  inline float kdtree_get_pt(const size_t idx, const size_t dim) const { return base[idx * stride_f + dim]; }
  template <class BBOX>
  bool kdtree_get_bbox(BBOX&) const { return false; }
};
// L2_Simple float == FLANN's L2_Simple<float> used by pcl::KdTreeFLANN (laserMapping.cpp:631-634,673,753)
using Tree3f = nanoflann::KDTreeSingleIndexAdaptor<nanoflann::L2_Simple_Adaptor<float, StridedCloud>, StridedCloud, 3,
                                                   int32_t>;
struct Handle {
  StridedCloud cloud;
  std::unique_ptr<Tree3f> tree;
};
}  // namespace

#define REF_API extern "C" __attribute__((visibility("default")))

REF_API void* ref_kdtree_build(const float* xyz, int n, int stride_bytes, int leaf_max) {
  Handle* h = new Handle{StridedCloud{xyz, (size_t)n, stride_bytes / 4}, nullptr};
  h->tree.reset(new Tree3f(3, h->cloud, nanoflann::KDTreeSingleIndexAdaptorParams(leaf_max)));
  h->tree->buildIndex();
  return h;
}
REF_API void ref_kdtree_free(void* hv) { delete static_cast<Handle*>(hv); }

REF_API void ref_kdtree_knn(void* hv, const float* q, int nq, int q_stride_bytes, int k, int32_t* idx, float* d2) {
  Handle* h = static_cast<Handle*>(hv);
  int qs = q_stride_bytes / 4;
  for (int i = 0; i < nq; ++i) {
    nanoflann::KNNResultSet<float, int32_t> rs(k);
    for (int j = 0; j < k; ++j) {
      idx[(size_t)i * k + j] = -1;
      d2[(size_t)i * k + j] = std::numeric_limits<float>::infinity();
    }
    rs.init(idx + (size_t)i * k, d2 + (size_t)i * k);
    h->tree->findNeighbors(rs, q + (size_t)i * qs, nanoflann::SearchParams());
  }
}

// ScanContext ring-key search exactly as Scancontext.cpp:270-295 sets it up:
// KDTreeVectorOfVectorsAdaptor<KeyMat,float> (metric_L2, dim 20, leaf 10), 10-NN.
REF_API void ref_ringkey_knn(const float* keys, int n, int dim, const float* query, int k, uint64_t* idx, float* d2) {
  using KeyMat = std::vector<std::vector<float>>;
  using InvKeyTree = KDTreeVectorOfVectorsAdaptor<KeyMat, float>;
  KeyMat mat(n, std::vector<float>(dim));
  for (int i = 0; i < n; ++i)
    for (int d = 0; d < dim; ++d) mat[i][d] = keys[(size_t)i * dim + d];
  InvKeyTree tree(dim, mat, 10);
  std::vector<size_t> ind(k);
  std::vector<float> dist(k);
  nanoflann::KNNResultSet<float> rs(k);
  rs.init(ind.data(), dist.data());
  tree.index->findNeighbors(rs, query, nanoflann::SearchParams(10));
  for (int j = 0; j < k; ++j) {
    idx[j] = j < (int)rs.size() ? ind[j] : (uint64_t)-1;
    d2[j] = j < (int)rs.size() ? dist[j] : std::numeric_limits<float>::infinity();
  }
}
