// ref_ikd.cpp -- thin C wrapper around the REFERENCE'S OWN ikd-Tree (/root/reference/src/ikd-Tree/ikd_Tree.{h,cpp}),
// compiled from where it lies after the five brace repairs of oracle/patches/ikd_tree_fix.py (the fork's annotation pass
// broke its block structure).  TEST INFRASTRUCTURE: it pins SURVEY §8 rows d1-d3 -- Build
// (ikd_Tree.cpp:471-492, 842-919), Nearest_Search (:495-547, 1381-1604) and Add_Points with down-sampling (:570-706) --
// the oracle restatement (ilsm_oracle.cpp: orc_knn_*, orc_ikd_add_points) is checked against it by
// tests/test_ref_ikd_cpu.py and tests/golden/ikd_reference.npz.  Output: oracle/_ref/libref_ikd.so (git-ignored).
// The tree is instantiated for pcl::PointXYZ as in the reference (mapOptimization.hpp:117, parameters.h_ouster:124);
// <pcl/point_types.h> and <Eigen/StdVector> come from oracle/shims/ (layout-identical stand-ins).
#include "ikd_Tree_fixed.cpp"  // generated into the temporary build directory by the Makefile, never stored here

#include <cstdint>

#define REF_API extern "C" __attribute__((visibility("default")))

namespace {
using Tree = KD_TREE<pcl::PointXYZ>;
using PV = Tree::PointVector;
PV to_pv(const float* xyz, int n, int stride_f) {
  PV v;
  v.reserve(n);
  for (int i = 0; i < n; ++i) v.emplace_back(xyz[(size_t)i * stride_f], xyz[(size_t)i * stride_f + 1], xyz[(size_t)i * stride_f + 2]);
  return v;
}
}  // namespace

// mapOptimization.cpp:504: `ikdtree(new KD_TREE<GroundPlanePointType>(0.3, 0.6, 0.4))` = (delete, balance, box length)
REF_API void* ref_ikd_create(float delete_param, float balance_param, float box_length) {
  return new Tree(delete_param, balance_param, box_length);
}
REF_API void ref_ikd_free(void* h) { delete static_cast<Tree*>(h); }
REF_API int ref_ikd_size(void* h) { return static_cast<Tree*>(h)->size(); }
// ikdtree->Build(points)  mapOptimization.cpp:192
REF_API void ref_ikd_build(void* h, const float* xyz, int n, int stride_bytes) {
  static_cast<Tree*>(h)->Build(to_pv(xyz, n, stride_bytes / 4));
}
// ikdtree->Nearest_Search(p, k, pts, d2)  mapOptimization.cpp:393 ; out_xyz: nq*k*3 floats, out_d2: nq*k floats,
// out_n: neighbours found per query (missing ones are left untouched)
REF_API void ref_ikd_nearest(void* h, const float* q_xyz, int nq, int stride_bytes, int k, float* out_xyz, float* out_d2,
                             int32_t* out_n) {
  Tree* t = static_cast<Tree*>(h);
  const int sf = stride_bytes / 4;
  for (int i = 0; i < nq; ++i) {
    PV pts;
    std::vector<float> d2;
    t->Nearest_Search(pcl::PointXYZ(q_xyz[(size_t)i * sf], q_xyz[(size_t)i * sf + 1], q_xyz[(size_t)i * sf + 2]), k, pts, d2);
    out_n[i] = (int32_t)pts.size();
    for (size_t j = 0; j < pts.size() && j < (size_t)k; ++j) {
      out_xyz[((size_t)i * k + j) * 3 + 0] = pts[j].x;
      out_xyz[((size_t)i * k + j) * 3 + 1] = pts[j].y;
      out_xyz[((size_t)i * k + j) * 3 + 2] = pts[j].z;
      out_d2[(size_t)i * k + j] = d2[j];
    }
  }
}
// ikdtree->Add_Points(points, downsample_on)  mapOptimization.cpp:475 ; returns the reference's own return value
REF_API int ref_ikd_add_points(void* h, const float* xyz, int n, int stride_bytes, int downsample_on) {
  PV v = to_pv(xyz, n, stride_bytes / 4);
  return static_cast<Tree*>(h)->Add_Points(v, downsample_on != 0);
}
// ikdtree->flatten(Root_Node, storage, NOT_RECORD)  mapOptimization.cpp:224 ; tree order (depends on the tree's shape)
REF_API int ref_ikd_flatten(void* h, float* out_xyz, int capacity) {
  Tree* t = static_cast<Tree*>(h);
  PV v;
  t->flatten(t->Root_Node, v, NOT_RECORD);
  const int n = (int)v.size();
  for (int i = 0; i < n && i < capacity; ++i) out_xyz[i * 3] = v[i].x, out_xyz[i * 3 + 1] = v[i].y, out_xyz[i * 3 + 2] = v[i].z;
  return n;
}
