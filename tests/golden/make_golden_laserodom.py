#!/usr/bin/env python
"""Generates tests/golden/laserodom_reference.npz: the residual blocks the REFERENCE's own scan-to-scan association builds
(/root/reference/src/laserOdometry.cpp:417-713 + TransformToStart :147-172, built into oracle/_ref/libref_laserodom.so by
oracle/Makefile through oracle/patches/laserodom_extract.py) for two consecutive seeded synthetic frames at two poses.
The feature clouds are regenerated from the seeds by the tests.  Run in the build container (needs /root/reference):
  python tests/golden/make_golden_laserodom.py"""
import functools
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
import oracle  # noqa: E402
import ilsm_b200 as ilsm  # noqa: E402

POSES = {"identity": np.array([0, 0, 0, 1, 0, 0, 0.0]),
         "moved": np.array([0.002, -0.001, 0.01, 0.99995, 0.24, 0.02, -0.01])}
POSES["moved"][:4] /= np.linalg.norm(POSES["moved"][:4])


@functools.lru_cache(maxsize=1)
def feature_clouds():
    """(last less-sharp, last less-flat, sharp, flat) of two consecutive frames, as scanRegistration emits them."""
    S = ilsm.synth
    scene = S.Scene()
    q0, t0 = S.default_pose()
    f0 = S.make_frame(scene, q0, t0, seed=0x5EED0D00)[0]
    q1 = S.quat_mul(q0, S.quat_from_rotvec([0, 0, 0.02]))
    f1 = S.make_frame(scene, q1, np.asarray(t0) + [0.25, 0.03, 0.0], seed=0x5EED0D01)[0]
    a, b = oracle.extract_features(f0), oracle.extract_features(f1)
    return a["cloud"][a["less_sharp_idx"]], a["less_flat"], b["cloud"][b["sharp_idx"]], b["cloud"][b["flat_idx"]]


def plane_normal_form(plane):
    """LidarPlaneFactor's (j, l, m) as the unit normal and offset the library reports: n = (j - l) x (j - m) / |.|, d = -j.n"""
    j, l, m = plane[:, 3:6], plane[:, 6:9], plane[:, 9:12]
    n = np.cross(j - l, j - m)
    n /= np.linalg.norm(n, axis=1, keepdims=True)
    return n, -(j * n).sum(1)


if __name__ == "__main__":
    lc, ls, sh, fl = feature_clouds()
    out = {}
    for name, qt in POSES.items():
        e, p, c = oracle.ref_odom_associate(lc, ls, sh, fl, qt)
        out[name + "/edge"], out[name + "/plane"], out[name + "/counters"] = e, p, np.array(c)
        print(name, e.shape, p.shape, c)
    np.savez_compressed(os.path.join(os.path.dirname(os.path.abspath(__file__)), "laserodom_reference.npz"), **out)
