#!/usr/bin/env python
"""Generates tests/golden/lasermapping_reference.npz: the state of the REFERENCE's own cube map
(/root/reference/src/laserMapping.cpp:327-623, 875-945, 984-1004, built into oracle/_ref/libref_lasermapping.so by
oracle/Makefile through oracle/patches/lasermapping_extract.py) after a seeded 61-frame walk that rolls the 21x21x11 window
along every axis in both directions: per frame the insertion pose, the window centre, the number of valid cubes and the
map / stack sizes; at the end the point count of every cube and one SHA-256 over all cube contents in index order.
Also the residual blocks of the reference's scan-to-map association (:624-873, recording ceres::Problem) on the config-1
shape: the 5-NN gate, the line / plane tests and point_a / point_b / unit normals as the reference code builds them.
Run in the build container (needs /root/reference):  python tests/golden/make_golden_lasermapping.py"""
import hashlib
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)

STEPS = [(60, 0, 0)] * 9 + [(0, -60, 0)] * 9 + [(0, 0, 40)] * 5 + [(-60, 0, 0)] * 12 + [(0, 60, 0)] * 12 + [(0, 0, -40)] * 8 + [(35, 25, -10)] * 6
N_CUBES = 21 * 21 * 11


def walk():
    """(corner cloud, surf cloud, odometry pose) per frame: sensor-frame points up to 70 m away, so every frame touches
    several 50 m cubes; 60 m steps move the window centre every frame."""
    rng = np.random.default_rng(5)
    t = np.zeros(3)
    for k, st in enumerate(STEPS):
        t = t + np.array(st, float) + rng.normal(0, 0.5, 3)
        ang = 0.02 * k
        qt = np.concatenate([[0, 0, np.sin(ang / 2), np.cos(ang / 2)], t])
        corner = np.concatenate([rng.uniform(-70, 70, (300, 3)), rng.integers(0, 64, (300, 1))], 1).astype(np.float32)
        surf = np.concatenate([rng.uniform(-70, 70, (1500, 3)) * [1, 1, 0.3], rng.integers(0, 64, (1500, 1))], 1).astype(np.float32)
        yield corner, surf, qt


def cube_state(cube_fn):
    """(counts (2, N_CUBES), sha256 over all cube contents in (kind, index) order) for a `cube(which, index)` accessor."""
    counts = np.zeros((2, N_CUBES), np.int32)
    h = hashlib.sha256()
    for which in (0, 1):
        for idx in range(N_CUBES):
            pts = np.ascontiguousarray(cube_fn(which, idx), np.float32)
            counts[which, idx] = len(pts)
            h.update(pts.tobytes())
    return counts, h.hexdigest()


def association_case():
    """Map clouds, stacks and the initial pose of the config-1 shape (20 k-point map) for the scan-to-map association."""
    import ilsm_b200 as ilsm
    c = ilsm.synth.config1(n_map=20_000)
    return c["map_corner"], c["map_surf"], c["corner"], c["surf"], np.concatenate([c["q0"], c["t0"]])


if __name__ == "__main__":
    import oracle
    edge, plane = oracle.ref_map_associate(*association_case())
    print("association blocks", edge.shape, plane.shape)
    ref = oracle.RefLaserMapping(0.4, 0.8)
    poses, cens, nvalid, sizes = [], [], [], []
    for corner, surf, qt in walk():
        q, cen, valid, sz = ref.frame(corner, surf, qt)
        poses.append(q), cens.append(cen.copy()), nvalid.append(len(valid)), sizes.append(sz.copy())
    counts, digest = cube_state(ref.cube)
    print("frames", len(poses), "occupied cubes", int((counts > 0).sum()), "points", int(counts.sum()), "centre changes",
          sum(tuple(a) != tuple(b) for a, b in zip(cens[1:], cens[:-1])))
    np.savez_compressed(os.path.join(os.path.dirname(os.path.abspath(__file__)), "lasermapping_reference.npz"), poses=np.array(poses),
                        cen=np.array(cens), n_valid=np.array(nvalid), sizes=np.array(sizes), counts=counts, cubes_sha256=digest,
                        assoc_edge=edge, assoc_plane=plane)
