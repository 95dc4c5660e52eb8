"""The oracle's rolling cube map against the REFERENCE's own code: src/laserMapping.cpp's window roll, 5x5x3 gather, stack
VoxelGrids, transformUpdate, insertion and per-cube VoxelGrid (:327-623, 875-945, 984-1004) cut out of process()
(oracle/patches/lasermapping_extract.py) and compiled into oracle/_ref/libref_lasermapping.so.  Both sides run without the
optimisation block, so the comparison is on the map logic: 61 frames, the window rolling along every axis in both
directions.  Live when the library exists, and against its committed outputs everywhere."""
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "tests", "golden"))
from make_golden_lasermapping import association_case, cube_state, walk  # noqa: E402

GOLD = np.load(os.path.join(ROOT, "tests", "golden", "lasermapping_reference.npz"))


def run_oracle(oracle_mod):
    m = oracle_mod.CubeMap(0.4, 0.8)
    m.set_solve(False)
    poses, cens, nvalid, sizes = [], [], [], []
    for corner, surf, qt in walk():
        q, _, st = m.frame(corner, surf, qt)
        poses.append(q), cens.append(tuple(st.cen)), nvalid.append(st.n_valid)
        sizes.append((st.n_map_corner, st.n_map_surf, st.n_stack_corner, st.n_stack_surf))
        assert st.ran_optimization == 0
    return m, np.array(poses), np.array(cens), np.array(nvalid), np.array(sizes)


def test_oracle_cube_map_equals_reference_golden(oracle_mod):
    m, poses, cens, nvalid, sizes = run_oracle(oracle_mod)
    assert np.array_equal(poses, GOLD["poses"]) and np.array_equal(cens, GOLD["cen"])
    assert np.array_equal(nvalid, GOLD["n_valid"]) and np.array_equal(sizes, GOLD["sizes"])
    assert sum(tuple(a) != tuple(b) for a, b in zip(cens[1:], cens[:-1])) >= 12  # the window did roll, on all three axes
    assert len({c[0] for c in map(tuple, cens)}) > 2 and len({c[1] for c in map(tuple, cens)}) > 2 and len({c[2] for c in map(tuple, cens)}) > 1
    counts, digest = cube_state(m.cube)
    assert np.array_equal(counts, GOLD["counts"]) and (counts > 0).sum() > 500
    assert digest == str(GOLD["cubes_sha256"])


def test_oracle_cube_map_equals_reference_live(oracle_mod):
    if oracle_mod.ref_lasermapping() is None:
        pytest.skip("oracle/_ref/libref_lasermapping.so not built (needs the reference tree)")
    ref = oracle_mod.RefLaserMapping(0.4, 0.8)
    m = oracle_mod.CubeMap(0.4, 0.8)
    m.set_solve(False)
    for k, (corner, surf, qt) in enumerate(walk()):
        rq, rcen, rvalid, rs = ref.frame(corner, surf, qt)
        oq, _, st = m.frame(corner, surf, qt)
        assert np.array_equal(rq, oq) and tuple(rcen) == tuple(st.cen) and len(rvalid) == st.n_valid, k
        assert tuple(rs) == (st.n_map_corner, st.n_map_surf, st.n_stack_corner, st.n_stack_surf), k
        if k == 30:  # every cube, mid-way (and at the end)
            assert cube_state(ref.cube)[1] == cube_state(m.cube)[1], k
    assert cube_state(ref.cube)[1] == cube_state(m.cube)[1]


def check_association(f, edge, plane, tol):
    fe, fp = f[f["type"] == 1], f[f["type"] == 2]
    assert len(fe) == len(edge) and len(fp) == len(plane) and len(edge) > 100 and len(plane) > 500
    assert np.array_equal(fe["p"], edge[:, 0:3]) and np.array_equal(fp["p"], plane[:, 0:3])  # the same stack points pass the tests
    assert np.abs(fe["a"] - edge[:, 3:6]).max() <= tol and np.abs(fe["b"] - edge[:, 6:9]).max() <= tol
    assert np.abs(fp["a"] - plane[:, 3:6]).max() <= tol and np.abs(fp["b"][:, 0] - plane[:, 6]).max() <= tol


def test_oracle_map_association_equals_reference_golden(oracle_mod):
    """laserMapping.cpp:624-873 compiled with a recording ceres::Problem: which stack points get a factor (5-NN gate
    d2[4] < 1, lambda_2 > 3 lambda_1, all five plane distances <= 0.2) and the factors themselves.  The reference's
    SelfAdjointEigenSolver / colPivHouseholderQr run on the oracle's kernels there, so the fits agree to the last bit."""
    mc, ms, c, s, qt = association_case()
    check_association(oracle_mod.associate(mc, ms, c, s, qt), GOLD["assoc_edge"], GOLD["assoc_plane"], 0.0)


def test_oracle_map_association_equals_reference_live(oracle_mod):
    if oracle_mod.ref_lasermapping() is None:
        pytest.skip("oracle/_ref/libref_lasermapping.so not built (needs the reference tree)")
    mc, ms, c, s, qt = association_case()
    rng = np.random.default_rng(11)
    for k in range(3):
        q = qt.copy()
        q[:4] += rng.normal(0, 0.003, 4)
        q[:4] /= np.linalg.norm(q[:4])
        q[4:] += rng.normal(0, 0.1, 3)
        e, p = oracle_mod.ref_map_associate(mc, ms, c, s, q)
        check_association(oracle_mod.associate(mc, ms, c, s, q), e, p, 0.0)
