"""GPU parity of the scan-to-scan odometry path (laserOdometry.cpp) against the CPU oracle."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def two_frames(ilsm, oracle_mod):
    S = ilsm.synth
    scene = S.Scene(S.SEED_MAP)
    q0, t0 = S.default_pose()
    dq = S.quat_from_rotvec([0.002, -0.001, 0.02])
    dt = np.array([0.18, 0.03, 0.005])
    q1 = S.quat_mul(q0, dq)
    t1 = t0 + S.quat_to_mat(q0) @ dt
    c0, _ = S.make_frame(scene, q0, t0, seed=1)
    c1, _ = S.make_frame(scene, q1, t1, seed=2)
    f0 = oracle_mod.extract_features(c0)
    f1 = oracle_mod.extract_features(c1)
    return dict(last_corner=f0["cloud"][f0["less_sharp_idx"]], last_surf=f0["less_flat"],
                sharp=f1["cloud"][f1["sharp_idx"]], flat=f1["cloud"][f1["flat_idx"]], dq=dq, dt=dt)


def _maps(ctx, d, cell=2.5):
    return ctx.new_map().set_input_cloud(d["last_corner"], cell), ctx.new_map().set_input_cloud(d["last_surf"], cell)


@pytest.mark.parametrize("pose", ["identity", "near_truth"])
def test_odometry_association_parity(ctx, oracle_mod, two_frames, pose):
    d = two_frames
    qt = np.array([0, 0, 0, 1, 0, 0, 0.0]) if pose == "identity" else np.concatenate([d["dq"], d["dt"]])
    mc, ms = _maps(ctx, d)
    got = ctx.odometry(mc, ms, d["sharp"], d["flat"], qt[:4], qt[4:], factors_only=True)
    want = oracle_mod.odom_associate(d["last_corner"], d["last_surf"], d["sharp"], d["flat"], qt)
    assert np.array_equal(got["type"], want["type"])
    assert (want["type"] == 1).sum() > 100 and (want["type"] == 2).sum() > 400
    assert np.array_equal(got["p"], want["p"])
    ed = want["type"] == 1
    # the two (three) chosen points are an index decision: the records must be identical, not merely close
    assert np.array_equal(got["a"][ed], want["a"][ed]) and np.array_equal(got["b"][ed], want["b"][ed])
    pl = want["type"] == 2
    assert np.allclose(got["a"][pl], want["a"][pl], rtol=0, atol=1e-12)
    assert np.allclose(got["b"][pl][:, 0], want["b"][pl][:, 0], rtol=1e-12, atol=1e-12)
    mc.close(), ms.close()


@pytest.mark.parametrize("pose", ["identity", "moved"])
def test_odometry_association_matches_reference_code_golden(ctx, pose):
    """The CUDA association against the residual blocks the REFERENCE's own code builds (laserOdometry.cpp:417-713 compiled
    from the reference tree with a recording ceres::Problem, tests/golden/make_golden_laserodom.py): same sharp / flat
    points paired, the edge factors' two points identical, the plane factors' (j, l, m) as the same unit normal + offset."""
    import os, sys
    sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden"))
    from make_golden_laserodom import POSES, feature_clouds, plane_normal_form
    gold = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "laserodom_reference.npz"))
    lc, ls, sh, fl = feature_clouds()
    qt = POSES[pose]
    mc, ms = ctx.new_map().set_input_cloud(lc, 1.0), ctx.new_map().set_input_cloud(ls, 1.0)
    got = ctx.odometry(mc, ms, sh, fl, qt[:4], qt[4:], factors_only=True)
    edge, plane = gold[pose + "/edge"], gold[pose + "/plane"]
    fe, fp = got[got["type"] == 1], got[got["type"] == 2]
    assert len(fe) == len(edge) and len(fp) == len(plane)
    assert np.array_equal(fe["p"], edge[:, 0:3]) and np.array_equal(fe["a"], edge[:, 3:6]) and np.array_equal(fe["b"], edge[:, 6:9])
    assert np.array_equal(fp["p"], plane[:, 0:3])
    n, d = plane_normal_form(plane)
    assert np.abs(fp["a"] - n).max() <= 1e-12 and np.abs(fp["b"][:, 0] - d).max() <= 1e-10
    mc.close(), ms.close()


def test_odometry_solve_parity(ctx, oracle_mod, ilsm, two_frames):
    d = two_frames
    mc, ms = _maps(ctx, d)
    q, t, rep = ctx.odometry(mc, ms, d["sharp"], d["flat"], [0, 0, 0, 1], [0, 0, 0])
    wx, wsum, wnf = oracle_mod.odometry(d["last_corner"], d["last_surf"], d["sharp"], d["flat"], np.array([0, 0, 0, 1, 0, 0, 0.0]))
    for p in range(2):
        assert rep.pass_[p].num_edge_factors == wnf[2 * p] and rep.pass_[p].num_plane_factors == wnf[2 * p + 1]
        assert rep.pass_[p].termination == wsum[p].termination and rep.pass_[p].iterations == wsum[p].iterations
        assert abs(rep.pass_[p].final_cost - wsum[p].final_cost) <= 1e-5 * wsum[p].final_cost
    assert np.linalg.norm(t - wx[4:]) < 1e-4 and ilsm.synth.quat_angle(q, wx[:4]) < 1e-4
    # and it recovers the true inter-frame motion
    assert np.linalg.norm(t - d["dt"]) < 0.03 and ilsm.synth.quat_angle(q, d["dq"]) < 2e-3
    mc.close(), ms.close()


def test_odometry_cell_size_does_not_change_result(ctx, two_frames):
    """The 1-NN is exact whatever the voxel size of the search structure."""
    d = two_frames
    outs = []
    for cell in (1.0, 2.5, 6.0):
        mc, ms = _maps(ctx, d, cell)
        outs.append(ctx.odometry(mc, ms, d["sharp"], d["flat"], [0, 0, 0, 1], [0, 0, 0]))
        mc.close(), ms.close()
    for q, t, _ in outs[1:]:
        assert np.array_equal(q, outs[0][0]) and np.array_equal(t, outs[0][1])
