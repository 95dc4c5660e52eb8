#!/usr/bin/env python
"""Generates the golden vectors under tests/golden/ (run in the build container, where /root/reference exists).

  python tests/golden/make_golden.py

knn_nanoflann.npz   outputs of the REFERENCE'S OWN vendored nanoflann 1.3.2 (/root/reference/include/nanoflann.hpp built
                    into oracle/_ref by oracle/Makefile, NANOFLANN_FIRST_MATCH) -- the one piece of the reference that runs
                    here: a 3000-point map, 400 queries (near, far, exact duplicates), k = 1 / 5 / 8.
ringkey_nanoflann.npz  the reference's ScanContext ring-key 10-NN set-up (KDTreeVectorOfVectorsAdaptor, Scancontext.cpp:270-295)
                    on 120 synthetic descriptors (regenerated from seeds, checksummed).
orb_match_cv2.npz   outputs of cv2.BFMatcher(NORM_HAMMING, crossCheck).match -- the very library call the reference makes
                    (intensity_feature_tracker.cpp:631-635) -- on 300 x 280 synthetic 256-bit descriptors with ties.
oracle_regression.npz  outputs of the CPU oracle (NOT of the reference: the reference cannot be built) on
                    a small registration / front-end / ScanContext case; they pin the oracle and the CUDA path against
                    silent drift, nothing more.
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)

import ilsm_b200 as ilsm  # noqa: E402
import oracle  # noqa: E402


def main():
    assert oracle.ref() is not None, "oracle/_ref not built (needs /root/reference)"
    S = ilsm.synth
    rng = np.random.default_rng(0x601D)
    c = S.config1(n_map=3000)
    m = np.concatenate([c["map_corner"], c["map_surf"]])[:, :3].astype(np.float32)
    m[100] = m[7]          # exact duplicates: the tie-break (lower index first) is part of the contract
    m[101] = m[7]
    q = np.concatenate([m[rng.integers(0, len(m), 250)] + rng.normal(0, 0.3, (250, 3)),
                        rng.uniform(-80, 80, (100, 3)), m[[7, 100, 55]], m[rng.integers(0, len(m), 47)]]).astype(np.float32)
    tree = oracle.RefKdTree(m)
    out = {"map": m, "queries": q}
    for k in (1, 5, 8):
        i, d = tree.knn(q, k)
        out[f"idx_k{k}"], out[f"d2_k{k}"] = i, d
    np.savez_compressed(os.path.join(HERE, "knn_nanoflann.npz"), **out)

    db = S.sc_database(120, seed=0x601E)
    qs, ids, shifts = S.sc_queries(db, 4, seed=0x601F)
    cands = []
    for j in range(len(qs)):
        _, _, _, cd = oracle.sc_detect_loop_reference(db.astype(np.float64), qs[j].astype(np.float64))
        cands.append(cd)
    # the descriptors are regenerated from the seeds by the test (synth.sc_database / sc_queries); a checksum guards them
    import hashlib
    digest = np.frombuffer(hashlib.sha256(db.tobytes() + qs.tobytes()).digest(), np.uint8)
    np.savez_compressed(os.path.join(HERE, "ringkey_nanoflann.npz"), seeds=np.array([120, 0x601E, 4, 0x601F]), sha256=digest,
                        candidates=np.stack(cands))

    cs = S.config1(n_map=6000)
    x, sums, nf = oracle.register_aloam(cs["map_corner"], cs["map_surf"], cs["corner"], cs["surf"],
                                        np.concatenate([cs["q0"], cs["t0"]]))
    f = oracle.extract_features(cs["cloud"])
    d10, i10, s10 = oracle.sc_topk(db.astype(np.float64), qs[0].astype(np.float64), 10)
    np.savez_compressed(os.path.join(HERE, "oracle_regression.npz"), pose=x, factors=np.asarray(nf),
                        term=np.array([s.termination for s in sums]), iters=np.array([s.iterations for s in sums]),
                        n_cloud=len(f["cloud"]), sharp_idx=f["sharp_idx"], less_sharp_idx=f["less_sharp_idx"],
                        flat_idx=f["flat_idx"], n_less_flat=len(f["less_flat"]),
                        label_hist=np.bincount(f["label"] + 1, minlength=4), sc_dist=d10, sc_id=i10, sc_shift=s10)
    # ORB matching: outputs of the real OpenCV (cv2) BFMatcher the reference calls (intensity_feature_tracker.cpp:631-635)
    import cv2
    a = rng.integers(0, 256, (300, 32), dtype=np.uint8)
    b = rng.integers(0, 256, (280, 32), dtype=np.uint8)
    b[:150] = a[50:200] ^ (rng.random((150, 32)) < 0.04).astype(np.uint8)  # true correspondences with a few flipped bits
    b[200] = b[3]     # duplicated train rows and query rows: first-minimum tie-breaks
    a[250] = a[60]
    og = {"cur": a, "prev": b, "opencv_version": np.array(cv2.__version__)}
    for cc in (0, 1):
        m = cv2.BFMatcher(cv2.NORM_HAMMING, bool(cc)).match(a, b)
        og[f"q_cc{cc}"] = np.array([x.queryIdx for x in m], np.int32)
        og[f"t_cc{cc}"] = np.array([x.trainIdx for x in m], np.int32)
        og[f"d_cc{cc}"] = np.array([x.distance for x in m], np.float32)
    np.savez_compressed(os.path.join(HERE, "orb_match_cv2.npz"), **og)
    for n in ("knn_nanoflann.npz", "ringkey_nanoflann.npz", "oracle_regression.npz", "orb_match_cv2.npz"):
        print(n, os.path.getsize(os.path.join(HERE, n)), "bytes")


if __name__ == "__main__":
    main()
