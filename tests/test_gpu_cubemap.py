"""GPU parity of the device-resident rolling cube map (laserMapping.cpp process()) against the CPU oracle."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def _compare_cubes(cm, ocm, indices):
    n_pts = 0
    for idx in indices:
        for which in (0, 1):
            g = cm.cube(which, int(idx))
            w = ocm.cube(which, int(idx))
            assert g.shape == w.shape, (idx, which, g.shape, w.shape)
            assert np.array_equal(g, w), (idx, which)
            n_pts += len(g)
    return n_pts


def _valid_indices(centre, cen=(10, 10, 5)):
    def cc(v, c):
        k = int((v + 25.0) / 50.0) + c
        return k - 1 if v + 25.0 < 0 else k
    cI, cJ, cK = cc(centre[0], cen[0]), cc(centre[1], cen[1]), cc(centre[2], cen[2])
    out = []
    for i in range(cI - 2, cI + 3):
        for j in range(cJ - 2, cJ + 3):
            for k in range(cK - 1, cK + 2):
                if 0 <= i < 21 and 0 <= j < 21 and 0 <= k < 11:
                    out.append(i + 21 * j + 441 * k)
    return out


def test_seed_and_frames_match_oracle(ctx, oracle_mod, ilsm, cfg_full):
    """Seed the map with the config-1 world map, then run three frames of a short trajectory through process():
    pose, factor counts, termination and every valid cube's content must match the oracle."""
    S = ilsm.synth
    c = cfg_full
    cm = ilsm.CubeMap(ctx, 0.4, 0.8)  # default slab: 16384 points per cube
    ocm = oracle_mod.CubeMap(0.4, 0.8)
    cm.insert_world(c["map_corner"], c["map_surf"], c["t_true"])
    ocm.insert_world(c["map_corner"], c["map_surf"], c["t_true"])
    valid = _valid_indices(c["t_true"])
    assert _compare_cubes(cm, ocm, valid) > 50_000
    scene = c["scene"]
    q, t = c["q_true"], c["t_true"]
    rng = np.random.default_rng(4)
    for k in range(3):
        q = S.quat_mul(q, S.quat_from_rotvec([0, 0, 0.03]))
        t = t + np.array([0.25, 0.05, 0.0])
        cloud, _ = S.make_frame(scene, q, t, seed=500 + k)
        f = oracle_mod.extract_features(cloud)
        corner_last, surf_last = f["cloud"][f["less_sharp_idx"]], f["less_flat"]
        # odometry pose = truth with a small drift
        qo = S.quat_mul(q, S.quat_from_rotvec(rng.normal(0, 0.004, 3)))
        to = t + rng.normal(0, 0.04, 3)
        gq, gt, rep, st = cm.frame(corner_last, surf_last, qo, to)
        wq, wsum, wst = ocm.frame(corner_last, surf_last, np.concatenate([qo, to]))
        assert st.ran_optimization == wst.ran_optimization == 1
        assert (st.n_map_corner, st.n_map_surf, st.n_stack_corner, st.n_stack_surf, st.n_valid) == \
               (wst.n_map_corner, wst.n_map_surf, wst.n_stack_corner, wst.n_stack_surf, wst.n_valid)
        for p in range(2):
            assert rep.pass_[p].termination == wsum[p].termination and rep.pass_[p].iterations == wsum[p].iterations
        assert np.linalg.norm(gt - wq[4:]) < 1e-4 and S.quat_angle(gq, wq[:4]) < 1e-4
        assert np.linalg.norm(gt - t) < 0.05  # and the mapping corrects the drifted odometry
        _compare_cubes(cm, ocm, valid)
    cm.close()


def test_window_roll_matches_oracle(ctx, oracle_mod, ilsm):
    """Moving the centre near the border rolls the pointer grid (laserMapping.cpp:341-565): cubes keep their world
    position, the ones pushed off the edge are recycled empty."""
    rng = np.random.default_rng(9)
    cm = ilsm.CubeMap(ctx, 0.4, 0.8, 2048)
    ocm = oracle_mod.CubeMap(0.4, 0.8)
    cen = [10, 10, 5]
    for centre in ([0, 0, 0], [380.0, 10.0, 0.0], [520.0, -390.0, 30.0], [-300.0, -420.0, -160.0], [0.0, 0.0, 0.0]):
        ctr = np.array(centre, np.float64)
        corner = (ctr + rng.uniform(-60, 60, (800, 3))).astype(np.float32)
        surf = (ctr + rng.uniform(-60, 60, (3000, 3))).astype(np.float32)
        cm.insert_world(corner, surf, ctr)
        ocm.insert_world(corner, surf, ctr)
        # the whole 21x21x11 window must agree (sampled: every cube that holds points in the oracle + a stride)
        for idx in range(0, 4851, 7):
            assert len(cm.cube(0, idx)) == len(ocm.cube(0, idx)) and len(cm.cube(1, idx)) == len(ocm.cube(1, idx))
    occupied = [i for i in range(4851) if len(ocm.cube(1, i))]
    assert len(occupied) > 20
    _compare_cubes(cm, ocm, occupied)
    cm.close()


def test_guard_skips_optimisation_on_empty_map(ctx, oracle_mod, ilsm, cfg_small):
    """First frame of a run: the map is empty, the guard of laserMapping.cpp:624 skips the solve, the stack is inserted
    at the predicted pose."""
    c = cfg_small
    cm = ilsm.CubeMap(ctx, 0.4, 0.8, 4096)
    ocm = oracle_mod.CubeMap(0.4, 0.8)
    gq, gt, rep, st = cm.frame(c["corner"], c["surf"], c["q_true"], c["t_true"])
    wq, wsum, wst = ocm.frame(c["corner"], c["surf"], np.concatenate([c["q_true"], c["t_true"]]))
    assert st.ran_optimization == 0 and wst.ran_optimization == 0
    assert np.array_equal(gq, c["q_true"]) and np.array_equal(gt, c["t_true"])
    _compare_cubes(cm, ocm, _valid_indices(c["t_true"]))
    cm.close()


def test_cube_map_matches_reference_code_golden(ctx, ilsm):
    """The CUDA cube map against the state of the REFERENCE's own code (laserMapping.cpp's window roll, gather, stack
    VoxelGrids, insertion and per-cube VoxelGrid compiled from the reference tree, tests/golden/make_golden_lasermapping.py)
    after a 61-frame walk that rolls the 21x21x11 window along every axis in both directions.  The optimisation block is
    switched off on both sides (min_corner_map beyond any map size), so this compares the map logic: poses, window centre,
    valid cubes, sizes per frame, and every cube of both maps bit for bit at the end."""
    import os, sys
    sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden"))
    from make_golden_lasermapping import cube_state, walk
    gold = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "lasermapping_reference.npz"))
    cm = ilsm.CubeMap(ctx, 0.4, 0.8, 4096)
    opts = ilsm.default_opts()
    opts.min_corner_map = opts.min_surf_map = 1 << 30
    for k, (corner, surf, qt) in enumerate(walk()):
        q, t, rep, st = cm.frame(corner, surf, qt[:4], qt[4:], opts)
        assert st.ran_optimization == 0 and st.flags == 0
        assert np.array_equal(np.concatenate([q, t]), gold["poses"][k]), k
        assert tuple(st.cen) == tuple(gold["cen"][k]) and st.n_valid == gold["n_valid"][k], k
        assert (st.n_map_corner, st.n_map_surf, st.n_stack_corner, st.n_stack_surf) == tuple(gold["sizes"][k]), k
    counts, digest = cube_state(cm.cube)
    assert np.array_equal(counts, gold["counts"])
    assert digest == str(gold["cubes_sha256"])
    cm.close()
