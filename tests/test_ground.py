"""Ground-plane extraction (image_handler.h_ouster:41-100): oracle sanity on CPU, GPU parity (-m gpu).
PCL's SACSegmentation is not in this image and its sampler is rand()-based, so parity is against the oracle's
restatement of the PCL 1.10 RANSAC loop with a declared sampler (unpinned)."""
import numpy as np
import pytest


def _scene_frame(ilsm, seed=0, tilt=None):
    S = ilsm.synth
    c = S.config1(n_map=20_000)
    cloud = c["cloud"].copy()
    if tilt is not None:  # rotate the whole frame: the ground normal leaves the 15-degree cone
        R = S.quat_to_mat(S.quat_from_rotvec(tilt))
        cloud[:, :3] = (cloud[:, :3].astype(np.float64) @ R.T).astype(np.float32)
    return cloud


def test_oracle_ground_plane_finds_the_floor(oracle_mod, ilsm):
    cloud = _scene_frame(ilsm)
    g, co, info = oracle_mod.ground_plane(cloud)
    assert info["accepted"] and info["n_band"] > 20_000 and info["n_best"] > 0.8 * info["n_band"]
    # the synthetic sensor sits 1.5 m above the z = 0 ground with a small roll/pitch (default_pose)
    assert abs(co[3] - 1.5) < 0.02 and co[2] > 0.999
    assert len(g) > 25_000 and np.all(g[:, 2] < 0)
    # different sampler seeds find the same floor
    g2, co2, _ = oracle_mod.ground_plane(cloud, seed=12345)
    assert np.allclose(co, co2, atol=2e-3) and abs(len(g) - len(g2)) < 0.02 * len(g)


def test_oracle_ground_plane_rejects_tilted_floor(oracle_mod, ilsm):
    cloud = _scene_frame(ilsm, tilt=[0.5, 0.0, 0.0])
    g, co, info = oracle_mod.ground_plane(cloud)
    assert not info["accepted"] and len(g) == 0


def test_oracle_ransac_replay_follows_adaptive_bound(oracle_mod, ilsm):
    cloud = _scene_frame(ilsm)
    _, _, info = oracle_mod.ground_plane(cloud)
    counts = info["counts"]
    it = info["iterations"]
    assert 1 <= it <= 51 and info["best"] < it
    assert counts[info["best"]] == max(counts[:it])  # best of the hypotheses the loop actually visited


# ------------------------------------------------------------------------------------------------- GPU parity
@pytest.mark.gpu
@pytest.mark.parametrize("case", ["plain", "pcl32", "tilted", "seed7", "tight"])
def test_gpu_ground_matches_oracle(ctx, oracle_mod, ilsm, case):
    kw = {}
    cloud = _scene_frame(ilsm, tilt=[0.5, 0.0, 0.0] if case == "tilted" else None)
    if case == "seed7":
        kw = dict(seed=7)
    if case == "tight":
        kw = dict(distance_threshold=0.004, band=0.01, max_iterations=20)
    arg = cloud
    if case == "pcl32":  # pcl::PointXYZI layout
        arg = np.zeros((len(cloud), 8), np.float32)
        arg[:, :3] = cloud[:, :3]
    ge = ilsm.GroundExtractor(ctx)
    g, co, info = ge.extract(arg, **kw)
    okw = {("dist_thresh" if k == "distance_threshold" else k): v for k, v in kw.items()}
    wg, wco, winfo = oracle_mod.ground_plane(cloud, **okw)
    assert info.n_band == winfo["n_band"]
    assert (info.best_hypothesis, info.n_best_inliers, info.iterations) == (winfo["best"], winfo["n_best"], winfo["iterations"])
    assert bool(info.accepted) == winfo["accepted"]
    assert np.allclose(co, wco, rtol=0, atol=1e-6)
    if winfo["accepted"]:
        assert len(g) > 1000
        if np.array_equal(co, wco):
            assert np.array_equal(g, wg)
        else:  # a coefficient landed on the other side of a float rounding: the sets may differ at the band edge
            assert abs(len(g) - len(wg)) <= 3
    else:
        assert len(g) == 0
    ge.close()


@pytest.mark.gpu
def test_gpu_ground_degenerate_inputs(ctx, ilsm):
    ge = ilsm.GroundExtractor(ctx)
    g, co, info = ge.extract(np.zeros((0, 4), np.float32))
    assert len(g) == 0 and info.n_band == 0
    high = np.random.default_rng(0).uniform(0.5, 3.0, (5000, 4)).astype(np.float32)  # nothing in the z band
    g, co, info = ge.extract(high)
    assert len(g) == 0 and info.n_band == 0 and info.best_hypothesis == -1
    ge.close()
