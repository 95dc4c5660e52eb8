"""GPU parity of the mapOptimization callback (the mapping node spot.launch starts: ground extraction -> VoxelGrid(0.8) ->
plane association against the ground map -> LM <= 10 -> CONVERGENCE-gated update -> ikd-Tree style Add_Points) against
the oracle composition, on a short sequence over open ground."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def _sequence(S, n, oracle_mod):
    scene = S.Scene(S.SEED_MAP)
    q0, t0 = S.default_pose()
    R0 = S.quat_to_mat(q0)
    rng = np.random.default_rng(11)
    out = []
    for k in range(n):
        q = S.quat_mul(q0, S.quat_from_rotvec([0, 0, 0.02 * k]))
        t = t0 + np.array([0.25 * k, 0.05 * k, 0.0])
        cloud, _ = S.make_frame(scene, q, t, seed=900 + k)
        f = oracle_mod.extract_features(cloud)
        qo = S.quat_mul(S.quat_mul(S.quat_inv(q0), q), S.quat_from_rotvec(rng.normal(0, 0.002, 3)))
        to = R0.T @ (t - t0) + rng.normal(0, 0.02, 3)
        out.append((cloud, f["less_flat"], qo, to, R0.T @ (t - t0)))
    return out


def test_mapopt_sequence_matches_oracle(ctx, oracle_mod, ilsm):
    S = ilsm.synth
    seq = _sequence(S, 5, oracle_mod)
    mo = ilsm.MapOptimization(ctx)
    omo = oracle_mod.MapOptimization()
    for k, (cloud, plane, qo, to, t_true) in enumerate(seq):
        gq, gt, st = mo.frame(cloud, plane, qo, to)
        wx, info = omo.frame(cloud, plane, qo, to)
        assert st.n_ground == info["n_ground"] and st.n_plane_in == len(plane)
        assert (st.ground.best_hypothesis, st.ground.n_best_inliers) == (info["ground"]["best"], info["ground"]["n_best"])
        assert st.ran_optimization == info["ran_optimization"] == (1 if k else 0)
        assert np.linalg.norm(gt - wx[4:]) < 1e-4 and S.quat_angle(gq, wx[:4]) < 1e-4, k
        assert st.map_size == info["map_size"] == len(mo), k
        if k:
            assert st.n_query == info["n_query"] > 1000
            assert bool(st.converged) == info["converged"] and st.solve.termination == info["summary"].termination
            assert st.solve.iterations == info["summary"].iterations
            assert st.solve.num_plane_factors == info["n_plane_factors"] > 1000 and st.solve.num_edge_factors == 0
            key = np.concatenate([np.array(st.q_key[:]), np.array(st.t_key[:])])
            assert np.allclose(key, info["key"], rtol=0, atol=1e-9)
            assert np.linalg.norm(gt - t_true) < 0.1  # the drifted odometry is corrected in height / roll / pitch
        # the ground map itself (flatten): same point set, same order
        assert np.array_equal(mo.map_points(), omo.map), k
    mo.close()


def test_mapopt_first_frame_builds_without_downsampling(ctx, oracle_mod, ilsm):
    S = ilsm.synth
    cloud, plane, qo, to, _ = _sequence(S, 1, oracle_mod)[0]
    mo = ilsm.MapOptimization(ctx)
    gq, gt, st = mo.frame(cloud, plane, qo, to)
    assert st.ran_optimization == 0 and st.map_size == st.n_ground + len(plane) > 16384
    assert np.array_equal(gq, qo) and np.array_equal(gt, to)  # q/t_wmap_wodom start as the identity
    mo.close()
