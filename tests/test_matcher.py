"""Intensity-image feature back end: brute-force Hamming matching with cross check + best-fraction
selection (intensity_feature_tracker.cpp:631-648) and the 3D-3D alignment p2p_calculateRandT (:880-928).
CPU: the oracle against the committed outputs of the real cv2.BFMatcher (and against cv2 itself when importable).
GPU (-m gpu): the CUDA path through the C ABI against the oracle and the golden vectors; integer work, bit-exact."""
import os

import numpy as np
import pytest

G = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def _descriptors(rng, n1, n2, n_true, flip=0.04):
    a = rng.integers(0, 256, (n1, 32), dtype=np.uint8)
    b = rng.integers(0, 256, (n2, 32), dtype=np.uint8)
    k = min(n_true, n1, n2)
    b[:k] = a[n1 - k:] ^ (rng.random((k, 32)) < flip).astype(np.uint8)
    if n2 > 10 and n1 > 10:
        b[n2 - 1] = b[2]
        a[1] = a[n1 - 1]
    return a, b


def test_oracle_matcher_matches_opencv_golden(oracle_mod):
    g = np.load(os.path.join(G, "orb_match_cv2.npz"))
    for cc in (0, 1):
        q, t, d = oracle_mod.bf_match_hamming(g["cur"], g["prev"], bool(cc))
        assert np.array_equal(q, g[f"q_cc{cc}"]) and np.array_equal(t, g[f"t_cc{cc}"]) and np.array_equal(d, g[f"d_cc{cc}"])
    assert len(g["q_cc1"]) < len(g["q_cc0"]) == 300


def test_oracle_matcher_matches_live_opencv(oracle_mod):
    cv2 = pytest.importorskip("cv2")
    rng = np.random.default_rng(3)
    for n1, n2, nt in ((64, 50, 30), (700, 900, 400), (1, 5, 1)):
        a, b = _descriptors(rng, n1, n2, nt)
        for cc in (False, True):
            m = cv2.BFMatcher(cv2.NORM_HAMMING, cc).match(a, b)
            q, t, d = oracle_mod.bf_match_hamming(a, b, cc)
            assert [x.queryIdx for x in m] == q.tolist() and [x.trainIdx for x in m] == t.tolist()
            assert [x.distance for x in m] == d.tolist()


def test_oracle_good_matches_selection(oracle_mod):
    q = np.arange(10, dtype=np.int32)
    t = np.arange(10, dtype=np.int32)[::-1].copy()
    d = np.array([5, 3, 3, 9, 1, 3, 7, 1, 8, 2], np.float32)
    gq, gt, gd = oracle_mod.good_matches(q, t, d, 0.3)      # 10 * 0.3 = 3 -> i < 3.0: three matches
    assert gq.tolist() == [4, 7, 9] and gd.tolist() == [1, 1, 2]
    gq, _, _ = oracle_mod.good_matches(q[:7], t[:7], d[:7], 0.3)  # 7 * 0.3 = 2.1 -> i = 0, 1, 2
    assert gq.tolist() == [4, 1, 2]


def test_oracle_align_points_recovers_transform(oracle_mod, ilsm):
    S = ilsm.synth
    rng = np.random.default_rng(5)
    src = rng.uniform(-10, 10, (300, 3)).astype(np.float32)
    q = S.quat_from_rotvec([0.02, -0.01, 0.05])
    t = np.array([0.3, -0.1, 0.05])
    dst = (src.astype(np.float64) @ S.quat_to_mat(q).T + t + rng.normal(0, 0.005, (300, 3))).astype(np.float32)
    dst[:12] += 2.0  # outliers: HuberLoss(0.1)
    x, sm = oracle_mod.align_points(src, dst)
    assert S.quat_angle(x[:4], q) < 2e-3 and np.linalg.norm(x[4:] - t) < 2e-2
    # the fixed point agrees with an independent robust solver
    from scipy.optimize import least_squares
    from scipy.spatial.transform import Rotation as R

    def fun(p):
        return ((R.from_rotvec(p[:3]).apply(src.astype(np.float64)) + p[3:]) - dst.astype(np.float64)).ravel()
    # Ceres' HuberLoss acts on the squared norm of each 3-row block; compare on the inlier set instead
    inl = np.arange(12, 300)

    def fun_in(p):
        return ((R.from_rotvec(p[:3]).apply(src[inl].astype(np.float64)) + p[3:]) - dst[inl].astype(np.float64)).ravel()
    ref = least_squares(fun_in, np.zeros(6)).x
    assert np.linalg.norm(ref[3:] - x[4:]) < 5e-3


# ------------------------------------------------------------------------------------------------- GPU parity
@pytest.mark.gpu
def test_gpu_orb_match_golden_and_oracle(ctx, oracle_mod):
    g = np.load(os.path.join(G, "orb_match_cv2.npz"))
    for cc in (0, 1):
        m, good = ctx.orb_match(g["cur"], g["prev"], bool(cc), 0.3)
        assert np.array_equal(m["queryIdx"], g[f"q_cc{cc}"]) and np.array_equal(m["trainIdx"], g[f"t_cc{cc}"])
        assert np.array_equal(m["distance"], g[f"d_cc{cc}"])
        gq, gt, gd = oracle_mod.good_matches(g[f"q_cc{cc}"], g[f"t_cc{cc}"], g[f"d_cc{cc}"], 0.3)
        assert np.array_equal(good["queryIdx"], gq) and np.array_equal(good["trainIdx"], gt) and np.array_equal(good["distance"], gd)


@pytest.mark.gpu
@pytest.mark.parametrize("n1,n2,nt,frac", [(2000, 2000, 1200, 0.3), (4000, 3500, 900, 0.2), (7, 900, 5, 0.3), (900, 3, 3, 1.0),
                                          (1, 1, 1, 0.3)])
def test_gpu_orb_match_sizes(ctx, oracle_mod, n1, n2, nt, frac):
    rng = np.random.default_rng(n1 + n2)
    a, b = _descriptors(rng, n1, n2, nt)
    for cc in (True, False):
        m, good = ctx.orb_match(a, b, cc, frac)
        q, t, d = oracle_mod.bf_match_hamming(a, b, cc)
        assert np.array_equal(m["queryIdx"], q) and np.array_equal(m["trainIdx"], t) and np.array_equal(m["distance"], d)
        gq, gt, gd = oracle_mod.good_matches(q, t, d, frac)
        assert np.array_equal(good["queryIdx"], gq) and np.array_equal(good["trainIdx"], gt)
    m, good = ctx.orb_match(a, np.zeros((0, 32), np.uint8))
    assert len(m) == 0 and len(good) == 0


@pytest.mark.gpu
def test_gpu_align_points_matches_oracle(ctx, oracle_mod, ilsm):
    S = ilsm.synth
    rng = np.random.default_rng(8)
    for n in (4, 60, 600):
        src = rng.uniform(-15, 15, (n, 3)).astype(np.float32)
        q = S.quat_from_rotvec(rng.normal(0, 0.03, 3))
        t = rng.normal(0, 0.3, 3)
        dst = (src.astype(np.float64) @ S.quat_to_mat(q).T + t + rng.normal(0, 0.01, (n, 3))).astype(np.float32)
        dst[: n // 20] += 1.5
        gq, gt, sm = ctx.align_points(src, dst)
        wx, wsm = oracle_mod.align_points(src, dst)
        assert np.linalg.norm(gt - wx[4:]) < 1e-4 and S.quat_angle(gq, wx[:4]) < 1e-4
        assert sm.termination == wsm.termination and sm.iterations == wsm.iterations
        assert sm.num_edge_factors == n
        assert abs(sm.final_cost - wsm.final_cost) <= 1e-5 * max(1.0, abs(wsm.final_cost))
