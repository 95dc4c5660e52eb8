#!/usr/bin/env python
"""Generates tests/golden/scancontext_reference.npz: outputs of the REFERENCE's own ScanContext code
(/root/reference/src/Scancontext.cpp + include/Scancontext.h compiled unmodified into oracle/_ref/libref_scancontext.so by
oracle/Makefile, with the vendored nanoflann): descriptors of three seeded frames, ring / sector keys, column-shift distances
and aligning shifts of 48 (query, keyframe) pairs, and detectLoopClosureID (tree query + candidate loop + threshold) for 12
queries over a 300-keyframe database.  Run in the build container (needs /root/reference):
  python tests/golden/make_golden_scancontext.py"""
import functools
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
import ilsm_b200 as ilsm  # noqa: E402


@functools.lru_cache(maxsize=1)
def frames():
    S = ilsm.synth
    scene = S.Scene()
    q0, t0 = S.default_pose()
    qc, tc = S.corridor_poses(3)[1]
    return (S.make_frame(scene, q0, t0, seed=0x5EED0E01)[0], S.make_frame(S.Scene(corridor=True, length=60.0), qc, tc, seed=0x5EED0E02)[0],
            S.make_frame(scene, q0, np.asarray(t0) + [3.0, -2.0, 0.0], seed=0x5EED0E03, fov_deg=22.0)[0])


@functools.lru_cache(maxsize=1)
def database():
    """300 keyframe descriptors, 12 queries (shifted, noisy copies of known keyframes; the last four are unrelated to the
    database: no loop), and the 48 (query, keyframe) pairs whose distance is recorded."""
    S = ilsm.synth
    db = S.sc_database_range(0, 300, 300)
    q, ids, shifts = S.sc_queries(db, 8, seed=3)
    other = S.sc_database_range(5000, 5004, 6000)
    queries = np.concatenate([q, other]).astype(np.float32)
    pairs = [(j, int(c)) for j in range(12) for c in (ids[j % 8], 5, 77, 250)]
    return db, queries, ids, shifts, pairs


if __name__ == "__main__":
    import oracle
    out = {}
    for k, cloud in enumerate(frames()):
        d = oracle.ref_sc_make(cloud)
        rk, sk = oracle.ref_sc_keys(d)
        out[f"desc{k}"], out[f"ring_key{k}"], out[f"sector_key{k}"] = d, rk, sk
    db, queries, ids, shifts, pairs = database()
    dist = [oracle.ref_sc_distance(queries[j], db[c]) for j, c in pairs]
    out["pair_dist"], out["pair_shift"] = np.array([d for d, _ in dist]), np.array([s for _, s in dist], np.int32)
    det = [oracle.ref_sc_detect(db, queries[j]) for j in range(len(queries))]
    out["detect_id"], out["detect_yaw"] = np.array([i for i, _ in det], np.int32), np.array([y for _, y in det], np.float32)
    print("detect", out["detect_id"], "true", ids)
    np.savez_compressed(os.path.join(os.path.dirname(os.path.abspath(__file__)), "scancontext_reference.npz"), **out)
