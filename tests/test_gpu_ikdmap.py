"""GPU parity of the ikd-Tree style incremental map (Build / Add_Points with down-sampling / Nearest_Search)
against the literal point-by-point CPU restatement."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def _rows(a):
    a = np.ascontiguousarray(a[:, :3], np.float32)
    return a[np.lexsort((a[:, 2], a[:, 1], a[:, 0]))]


def test_add_points_downsample_matches_sequential_reference(ctx, oracle_mod):
    rng = np.random.default_rng(5)
    # Build is NOT down-sampled (mapOptimization.cpp:192): several points per 0.4 m box survive until touched
    base = rng.uniform(-6, 6, (3000, 3)).astype(np.float32)
    base[:, 2] *= 0.05
    m = ctx.new_map().set_input_cloud(base)
    ref = base.copy()
    for frame in range(4):
        add = (rng.uniform(-7, 7, (1500, 3)) * [1, 1, 0.05]).astype(np.float32)
        add[:200] = add[200:400] + rng.normal(0, 0.01, (200, 3)).astype(np.float32)  # several new points per box
        m.add_points(add, True, 0.4)
        ref = oracle_mod.ikd_add_points(ref, add, 0.4, True)
        got = m.points()
        assert len(got) == len(ref)
        assert np.array_equal(_rows(got), _rows(ref))
    # invariant after the boxes have been touched: at most one point per touched 0.4 m box
    # and the search structure was rebuilt over the new point set
    q = (rng.uniform(-6, 6, (500, 3)) * [1, 1, 0.05]).astype(np.float32)
    idx, d2 = m.nearest_k_search(q, 5)
    pts = m.points()
    ri, rd = oracle_mod.knn_kdtree(pts[:, :3], q, 5)
    assert np.array_equal(d2, rd) and np.array_equal(idx, ri)
    m.close()


def test_add_points_ties_and_edges(ctx, oracle_mod):
    ds = 0.4
    c = np.array([0.2, 0.2, 0.2], np.float32)  # centre of box [0, 0.4)^3
    existing = np.array([c + [0.1, 0, 0]], np.float32)
    m = ctx.new_map().set_input_cloud(existing)
    add = np.array([c + [-0.1, 0, 0],      # same distance as the existing point: the NEW point wins the tie
                    c + [0, 0.1, 0],       # same distance again: the LATER new point wins
                    [5.0, 5.0, 5.0],       # empty box: simply added
                    [0.39999998, 0.1, 0.1]], np.float32)
    m.add_points(add, True, ds)
    ref = oracle_mod.ikd_add_points(existing, add, ds, True)
    assert np.array_equal(_rows(m.points()), _rows(ref))
    assert len(ref) == 2
    # append policy (downsample off) keeps everything
    m.add_points(add, False)
    assert len(m.points()) == 2 + len(add)
    # Add_Points on a never-built map behaves like Build + insert
    m2 = ctx.new_map()
    m2.add_points(add, True, ds)
    assert np.array_equal(_rows(m2.points()), _rows(oracle_mod.ikd_add_points(np.zeros((0, 3), np.float32), add, ds, True)))
    m.close(), m2.close()


def test_mapoptimization_style_loop(ctx, oracle_mod, ilsm, cfg_small):
    """mapOptimization.cpp:173-196,368-479 in miniature: Build from the first frame, then per frame plane-only
    registration (max 10 iterations) against the ikd-style map followed by Add_Points(0.4) of the registered points."""
    c = cfg_small
    rng = np.random.default_rng(1)
    surf = c["surf"]
    R = ilsm.synth.quat_to_mat(c["q_true"])
    world = (surf[:, :3].astype(np.float64) @ R.T + c["t_true"]).astype(np.float32)
    m = ctx.new_map().set_input_cloud(c["map_surf"])
    dummy = ctx.new_map().set_input_cloud(np.zeros((0, 3), np.float32))
    opts = ilsm.default_opts(outer_iterations=1, max_num_iterations=10, min_corner_map=0, min_surf_map=0)
    q, t, rep = ctx.register(dummy, m, np.zeros((0, 4), np.float32), surf, c["q0"], c["t0"], opts)
    assert rep.pass_[0].termination in (0, 1) and np.linalg.norm(t - c["t_true"]) < 0.05
    before = len(m.points())
    m.add_points(world, True, 0.4)
    ref = oracle_mod.ikd_add_points(c["map_surf"], world, 0.4, True)
    assert len(m.points()) == len(ref) and np.array_equal(_rows(m.points()), _rows(ref))
    assert len(ref) != before
    m.close(), dummy.close()
