#!/usr/bin/env python
"""Generates tests/golden/ikd_reference.npz from the REFERENCE'S OWN ikd-Tree (run in the build container, where
/root/reference exists):

  python tests/golden/make_golden_ikd.py

The reference's src/ikd-Tree/ikd_Tree.cpp is compiled from where it lies into oracle/_ref/libref_ikd.so (oracle/Makefile,
after the repairs of oracle/patches/ikd_tree_fix.py) and driven the way mapOptimization.cpp drives it:
KD_TREE<pcl::PointXYZ>(0.3, 0.6, 0.4) (:504), Build (:192), Nearest_Search(p, 5) (:393), Add_Points(points, true) (:475).
Stored: the inputs, the neighbour points and squared distances of Nearest_Search (k = 5 and k = 1) after Build and after
the last insertion, and the point set of the tree (flatten, rows sorted) after each Add_Points batch.  A second case holds
hand-made ties and box edges.  These vectors pin SURVEY section 8 rows d1-d3 for the oracle and for the CUDA path.
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)

import oracle  # noqa: E402


def rows(a):
    a = np.ascontiguousarray(a[:, :3], np.float32)
    return a[np.lexsort((a[:, 2], a[:, 1], a[:, 0]))]


def main():
    assert oracle.ref_ikd() is not None, "oracle/_ref/libref_ikd.so not built (needs /root/reference)"
    rng = np.random.default_rng(0x1CD7)
    # a ground-like slab with a wall: several points per 0.4 m box (Build does not down-sample)
    base = np.concatenate([rng.uniform(-8, 8, (2600, 3)) * [1, 1, 0.04], np.c_[np.full(400, 8.0), rng.uniform(-8, 8, 400), rng.uniform(0, 3, 400)]])
    base = base.astype(np.float32)
    base[50] = base[9]  # exact duplicates in the Build cloud
    q = np.concatenate([base[rng.integers(0, len(base), 220)] + rng.normal(0, 0.25, (220, 3)), rng.uniform(-30, 30, (60, 3)),
                        base[[9, 50, 77]]]).astype(np.float32)
    t = oracle.RefIkdTree(0.3, 0.6, 0.4).build(base)
    out = {"base": base, "queries": q}
    for k in (1, 5):
        p, d, c = t.nearest(q, k)
        out[f"build_pts_k{k}"], out[f"build_d2_k{k}"], out[f"build_cnt_k{k}"] = p, d, c
    for b in range(3):
        add = (rng.uniform(-10, 10, (1400, 3)) * [1, 1, 0.04]).astype(np.float32)
        add[:150] = add[150:300] + rng.normal(0, 0.01, (150, 3)).astype(np.float32)  # several new points per box
        add[300] = add[301]                                                            # duplicate new points
        out[f"add{b}"] = add
        out[f"ret{b}"] = np.int32(t.add_points(add, True))
        out[f"set{b}"] = rows(t.points())
    p, d, c = t.nearest(q, 5)
    out["final_pts_k5"], out["final_d2_k5"], out["final_cnt_k5"] = p, d, c
    t.close()
    # ties and box edges (the policy of ikd_Tree.cpp:617-637: strict <, new point seeds min_dist, box = [min, max))
    ce = np.array([0.2, 0.2, 0.2], np.float32)
    existing = np.array([ce + [0.1, 0, 0], [3.0, 3.0, 3.0], [3.1, 3.1, 3.1]], np.float32)
    add = np.array([ce + [-0.1, 0, 0], ce + [0, 0.1, 0], [5.0, 5.0, 5.0], [0.39999998, 0.1, 0.1], [0.4, 0.1, 0.1], [-0.0, -1e-9, 0.3],
                    [3.39, 3.3, 3.3], [3.2, 3.2, 3.2]], np.float32)
    t = oracle.RefIkdTree(0.3, 0.6, 0.4).build(existing)
    out["tie_existing"], out["tie_add"] = existing, add
    out["tie_ret"] = np.int32(t.add_points(add, True))
    out["tie_set"] = rows(t.points())
    t.close()
    path = os.path.join(HERE, "ikd_reference.npz")
    np.savez_compressed(path, **out)
    print("ikd_reference.npz", os.path.getsize(path), "bytes;", {k: (v.shape if hasattr(v, "shape") else v) for k, v in out.items() if k.startswith(("set", "ret", "tie_"))})


if __name__ == "__main__":
    main()
