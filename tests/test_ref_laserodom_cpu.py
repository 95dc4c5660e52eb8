"""The oracle's scan-to-scan data association against the REFERENCE's own code: src/laserOdometry.cpp:417-713 (closest
point, +-2.5-ring walks, distance gate, TransformToStart) cut out of the node's spin loop
(oracle/patches/laserodom_extract.py) and compiled into oracle/_ref/libref_laserodom.so with a ceres::Problem that records
the residual blocks.  Live when that library exists, and against its committed outputs everywhere."""
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "tests", "golden"))
from make_golden_laserodom import POSES, feature_clouds, plane_normal_form  # noqa: E402

GOLD = np.load(os.path.join(ROOT, "tests", "golden", "laserodom_reference.npz"))


def check(f, edge, plane):
    fe, fp = f[f["type"] == 1], f[f["type"] == 2]
    assert len(fe) == len(edge) and len(fp) == len(plane) and len(edge) > 100 and len(plane) > 400
    # the points are floats widened to double: exact
    assert np.array_equal(fe["p"], edge[:, 0:3]) and np.array_equal(fe["a"], edge[:, 3:6]) and np.array_equal(fe["b"], edge[:, 6:9])
    assert np.array_equal(fp["p"], plane[:, 0:3])
    n, d = plane_normal_form(plane)
    assert np.abs(fp["a"] - n).max() <= 1e-12 and np.abs(fp["b"][:, 0] - d).max() <= 1e-10


@pytest.mark.parametrize("pose", sorted(POSES))
def test_oracle_association_equals_reference_golden(oracle_mod, pose):
    lc, ls, sh, fl = feature_clouds()
    check(oracle_mod.odom_associate(lc, ls, sh, fl, POSES[pose]), GOLD[pose + "/edge"], GOLD[pose + "/plane"])
    assert tuple(GOLD[pose + "/counters"]) == (len(GOLD[pose + "/edge"]), len(GOLD[pose + "/plane"]))


def test_oracle_association_equals_reference_live(oracle_mod):
    if oracle_mod.ref_laserodom() is None:
        pytest.skip("oracle/_ref/libref_laserodom.so not built (needs the reference tree)")
    lc, ls, sh, fl = feature_clouds()
    rng = np.random.default_rng(7)
    for k in range(4):
        qt = np.concatenate([rng.normal(0, 0.01, 3), [1.0], rng.normal(0, 0.3, 3)])
        qt[:4] /= np.linalg.norm(qt[:4])
        e, p, c = oracle_mod.ref_odom_associate(lc, ls, sh, fl, qt)
        assert c == (len(e), len(p))
        check(oracle_mod.odom_associate(lc, ls, sh, fl, qt), e, p)
    # swapped roles (a different, larger "last" cloud) and a pose far enough that many points fail the 5 m gate
    qt = np.array([0, 0, 0.05, 1.0, 3.5, -2.0, 0.3])
    qt[:4] /= np.linalg.norm(qt[:4])
    e, p, _ = oracle_mod.ref_odom_associate(lc, ls, sh, fl, qt)
    f = oracle_mod.odom_associate(lc, ls, sh, fl, qt)
    assert (f["type"] == 1).sum() == len(e) and (f["type"] == 2).sum() == len(p) and (f["type"] == 0).sum() > 0
