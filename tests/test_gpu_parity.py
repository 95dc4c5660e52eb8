"""GPU parity tests (run on the B200 with -m gpu): the CUDA path, called through the C ABI, against the CPU oracle
on the same seeded inputs.  Bars (BASELINE.json north_star): k-NN indices bit-exact under the (d2, index)
tie-break; residual quantities within 1e-5 relative; converged pose within 1e-4 m / 1e-4 rad."""
import numpy as np
import pytest

from conftest import pose7

pytestmark = pytest.mark.gpu

REL = 1e-5        # residual / normal-equation tolerance stated by the north star
POSE_M = 1e-4     # metres
POSE_RAD = 1e-4   # radians


def _queries(rng, m, n, spread):
    return (m[rng.integers(0, len(m), n)] + rng.normal(0, spread, (n, 3))).astype(np.float32)


@pytest.mark.parametrize("k", [1, 5, 8])
@pytest.mark.parametrize("cell", [0.0, 0.5, 2.0])
def test_knn_bit_exact(ctx, oracle_mod, cfg_small, k, cell):
    rng = np.random.default_rng(100 + k)
    m = cfg_small["map_surf"]
    q = np.concatenate([_queries(rng, m, 1500, 0.3), _queries(rng, m, 300, 3.0),
                        rng.uniform(-150, 150, (60, 3)).astype(np.float32)])  # far queries: ring expansion + brute force
    lm = ctx.new_map().set_input_cloud(m, cell)
    assert len(lm) == len(m)
    idx, d2 = lm.nearest_k_search(q, k)
    ri, rd = oracle_mod.knn_kdtree(m, q, k)
    assert np.array_equal(d2, rd)
    assert np.array_equal(idx, ri)
    lm.close()


@pytest.mark.parametrize("k,cell,max_dist", [(5, 0.0, 0.0), (5, 0.5, 0.0), (1, 0.0, 0.0), (5, 0.0, 1.0), (8, 2.0, 0.0)])
def test_knn_large_query_set(ctx, oracle_mod, ilsm, cfg_full, k, cell, max_dist):
    """A whole organised frame as the query set (grid-stride path, no-returns included) plus far and out-of-map queries;
    small and large launches agree."""
    S = ilsm.synth
    c = cfg_full
    rng = np.random.default_rng(55 + k)
    m = c["map_surf"]
    R = S.quat_to_mat(c["q_true"])
    frame_world = (c["cloud"][:, :3].astype(np.float64) @ R.T + c["t_true"]).astype(np.float32)  # 65536 coherent queries
    far = rng.uniform(-400, 400, (500, 3)).astype(np.float32)
    q = np.concatenate([frame_world, far, _queries(rng, m, 3000, 2.0)])
    lm = ctx.new_map().set_input_cloud(m, cell)
    idx, d2 = lm.nearest_k_search(q, k, max_dist=max_dist)
    sel = np.concatenate([rng.choice(65536, 3000, replace=False), np.arange(65536, len(q))])
    ri, rd = oracle_mod.knn_kdtree(m, q[sel], k)
    if max_dist > 0:
        inside = rd < max_dist * max_dist
        assert inside.any() and (~inside).any()
        assert np.array_equal(idx[sel][inside], ri[inside]) and np.array_equal(d2[sel][inside], rd[inside])
    else:
        assert np.array_equal(d2[sel], rd)
        assert np.array_equal(idx[sel], ri)
    # a small launch (one resident wave) gives the same answers
    i2, dd2 = lm.nearest_k_search(q[:4000], k, max_dist=max_dist)
    if max_dist > 0:
        ins = dd2 < max_dist * max_dist
        assert np.array_equal(i2[ins], idx[:4000][ins])
    else:
        assert np.array_equal(i2, idx[:4000]) and np.array_equal(dd2, d2[:4000])
    lm.close()


def test_knn_strided_pcl_points_and_max_dist(ctx, oracle_mod, cfg_small):
    """pcl::PointXYZI layout (32-byte stride) for both map and queries; bounded search is exact inside max_dist."""
    rng = np.random.default_rng(7)
    m3 = cfg_small["map_corner"]
    m = np.zeros((len(m3), 8), np.float32)
    m[:, :3] = m3
    m[:, 4] = rng.uniform(0, 255, len(m3))
    q3 = _queries(rng, m3, 2000, 0.5)
    q = np.zeros((len(q3), 8), np.float32)
    q[:, :3] = q3
    lm = ctx.new_map().set_input_cloud(m)
    idx, d2 = lm.nearest_k_search(q, 5, max_dist=1.0)
    ri, rd = oracle_mod.knn_kdtree(m3, q3, 5)
    inside = rd < 1.0
    assert inside.any() and (~inside).any()
    assert np.array_equal(idx[inside], ri[inside]) and np.array_equal(d2[inside], rd[inside])
    full = inside.all(axis=1)
    assert np.array_equal(idx[full], ri[full])
    lm.close()


def test_knn_edge_cases(ctx, oracle_mod):
    # fewer map points than k, duplicates (exact ties -> lower index first), empty query set, empty map
    m = np.array([[0, 0, 0], [1, 0, 0], [1, 0, 0], [0, 2, 0]], np.float32)
    lm = ctx.new_map().set_input_cloud(m)
    idx, d2 = lm.nearest_k_search(np.array([[0.9, 0, 0], [50, 50, 50]], np.float32), 5)
    ri, rd = oracle_mod.knn_brute(m, np.array([[0.9, 0, 0], [50, 50, 50]], np.float32), 5)
    assert np.array_equal(idx, ri) and np.array_equal(d2, rd)
    assert idx[0].tolist() == [1, 2, 0, 3, -1]
    idx, d2 = lm.nearest_k_search(np.zeros((0, 3), np.float32), 5)
    assert idx.shape == (0, 5)
    lm.set_input_cloud(np.zeros((0, 3), np.float32))
    idx, d2 = lm.nearest_k_search(np.zeros((3, 3), np.float32), 5)
    assert (idx == -1).all() and np.isinf(d2).all()
    # symmetric ties on a lattice: six neighbours at distance 1 -> indices 0..4
    m = np.array([[1, 0, 0], [0, 1, 0], [-1, 0, 0], [0, -1, 0], [0, 0, 1], [0, 0, -1], [5, 5, 5]], np.float32)
    lm.set_input_cloud(m)
    idx, d2 = lm.nearest_k_search(np.zeros((1, 3), np.float32), 5)
    assert idx[0].tolist() == [0, 1, 2, 3, 4] and (d2 == 1.0).all()
    # NaN points are skipped at build time (mapOptimization.cpp:151 strips them first)
    m = np.array([[0, 0, 0], [np.nan, 0, 0], [2, 0, 0]], np.float32)
    lm.set_input_cloud(m)
    idx, d2 = lm.nearest_k_search(np.array([[1.2, 0, 0]], np.float32), 2)
    assert idx[0].tolist() == [2, 0]
    lm.close()


def test_knn_rebuild_reuses_handle(ctx, oracle_mod, cfg_small):
    """setInputCloud is called every frame on the same kd-tree object (laserMapping.cpp:631-634)."""
    rng = np.random.default_rng(9)
    lm = ctx.new_map()
    for n in (5000, 12000, 800):
        m = cfg_small["map_surf"][rng.permutation(len(cfg_small["map_surf"]))[:n]]
        q = _queries(rng, m, 500, 0.4)
        lm.set_input_cloud(m)
        idx, d2 = lm.nearest_k_search(q, 5)
        ri, rd = oracle_mod.knn_kdtree(m, q, 5)
        assert np.array_equal(idx, ri) and np.array_equal(d2, rd)
    lm.close()


def _maps(ctx, c):
    return ctx.new_map().set_input_cloud(c["map_corner"]), ctx.new_map().set_input_cloud(c["map_surf"])


def _cmp_factors(got, want):
    assert np.array_equal(got["type"], want["type"])
    assert np.array_equal(got["src"], want["src"])
    assert np.array_equal(got["p"], want["p"])
    pl = want["type"] == 2
    assert np.allclose(got["a"][pl], want["a"][pl], rtol=0, atol=1e-9)
    assert np.allclose(got["b"][pl][:, 0], want["b"][pl][:, 0], rtol=1e-9, atol=1e-9)
    ed = want["type"] == 1
    # the sign of an eigenvector is arbitrary: (a, b) may come out swapped (same residual up to sign)
    ga, gb, wa, wb = got["a"][ed], got["b"][ed], want["a"][ed], want["b"][ed]
    same = np.maximum(np.abs(ga - wa).max(1), np.abs(gb - wb).max(1))
    swap = np.maximum(np.abs(ga - wb).max(1), np.abs(gb - wa).max(1))
    assert (np.minimum(same, swap) < 1e-9).all()


def test_associate_parity(ctx, oracle_mod, cfg_small):
    c = cfg_small
    mc, ms = _maps(ctx, c)
    qt = pose7(c["q0"], c["t0"])
    fac, idx, d2 = ctx.associate(mc, ms, c["corner"], c["surf"], c["q0"], c["t0"], want_knn=True)
    want = oracle_mod.associate(c["map_corner"], c["map_surf"], c["corner"], c["surf"], qt)
    # the k-NN behind the factors: bit-exact wherever the 5th neighbour passes the d2 < 1 gate
    pw_c = oracle_mod.transform_points(qt, c["corner"])
    pw_s = oracle_mod.transform_points(qt, c["surf"])
    ri = np.concatenate([oracle_mod.knn_kdtree(c["map_corner"], pw_c, 5)[0], oracle_mod.knn_kdtree(c["map_surf"], pw_s, 5)[0]])
    rd = np.concatenate([oracle_mod.knn_kdtree(c["map_corner"], pw_c, 5)[1], oracle_mod.knn_kdtree(c["map_surf"], pw_s, 5)[1]])
    gate = rd[:, 4] < 1.0
    assert gate.sum() > 500
    assert np.array_equal(idx[gate], ri[gate]) and np.array_equal(d2[gate], rd[gate])
    assert (d2[~gate][:, 4] >= 1.0).all() or np.isinf(d2[~gate][:, 4]).any()
    _cmp_factors(fac, want)
    assert (fac["type"] == 1).sum() > 50 and (fac["type"] == 2).sum() > 300
    mc.close(), ms.close()


@pytest.mark.parametrize("copies", [3, 14])
def test_associate_parity_stacks_larger_than_one_wave(ctx, oracle_mod, cfg_small, copies):
    """Stacks with more points than the launch has warps (2960 on a B200): the kernel gathers chunks of up to 32 points per
    block and fits them on the lanes of one warp; 14 copies exceed 8 points per warp, so a block walks several chunks and
    one chunk straddles the corner / surf boundary.  Same factors as the oracle, point by point."""
    c = cfg_small
    rng = np.random.default_rng(17 + copies)
    def tile(a):
        b = np.tile(a, (copies, 1)).astype(np.float32)
        b[:, :3] += rng.normal(0.0, 0.02, (len(b), 3)).astype(np.float32)
        return b
    corner, surf = tile(c["corner"]), tile(c["surf"])
    assert len(corner) + len(surf) > 2960 * (8 if copies > 8 else 1)
    mc, ms = _maps(ctx, c)
    qt = pose7(c["q0"], c["t0"])
    fac = ctx.associate(mc, ms, corner, surf, c["q0"], c["t0"])
    fac = fac[0] if isinstance(fac, tuple) else fac
    want = oracle_mod.associate(c["map_corner"], c["map_surf"], corner, surf, qt)
    _cmp_factors(fac, want)
    assert (fac["type"] == 1).sum() > 50 * copies and (fac["type"] == 2).sum() > 300 * copies
    mc.close(), ms.close()


@pytest.mark.parametrize("line_res,plane_res", [(0.4, 0.8), (0.1, 0.2)])
def test_register_frame_parity(ctx, oracle_mod, ilsm, cfg_full, line_res, plane_res):
    """ilsm_register_frame (front end -> VoxelGrid stacks -> association / solve with the stack sizes read on the device)
    against the chained oracle.  With the fine leaves the stacks (about 9 k points) exceed the warps of the association
    launch, which only knows the pre-filter sizes: its one-point-per-warp kernel then takes several rounds."""
    c = cfg_full
    frame = np.ascontiguousarray(c["cloud"][:, :4], np.float32)
    mc, ms = _maps(ctx, c)
    q, t, rep, sizes = ctx.register_frame(mc, ms, frame, c["q0"], c["t0"], 0.3, line_res, plane_res)
    f = oracle_mod.extract_features(frame)
    sc_ = oracle_mod.voxelgrid(f["cloud"][f["less_sharp_idx"]], line_res)
    ss_ = oracle_mod.voxelgrid(f["less_flat"], plane_res)
    assert sizes == (len(f["less_sharp_idx"]), len(f["less_flat"]), len(sc_), len(ss_))
    if line_res < 0.4:
        assert len(sc_) + len(ss_) > 2960
    x, _, _ = oracle_mod.register_aloam(c["map_corner"], c["map_surf"], sc_, ss_, pose7(c["q0"], c["t0"]))
    assert np.linalg.norm(t - x[4:]) < 1e-6 and ilsm.synth.quat_angle(q, x[:4]) < 1e-6
    mc.close(), ms.close()


def test_associate_matches_reference_code_golden(ctx):
    """The CUDA association against the residual blocks the REFERENCE's own code builds (laserMapping.cpp:624-873 compiled
    from the reference tree with a recording ceres::Problem, tests/golden/make_golden_lasermapping.py): the same stack points
    pass the 5-NN gate and the line / plane tests, and point_a / point_b / unit normal / offset agree (the fits run on
    different eigen / QR kernels: 1e-9)."""
    import os, sys
    sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden"))
    from make_golden_lasermapping import association_case
    gold = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "lasermapping_reference.npz"))
    mcn, msn, c, s, qt = association_case()
    mc, ms = ctx.new_map().set_input_cloud(mcn), ctx.new_map().set_input_cloud(msn)
    fac = ctx.associate(mc, ms, c, s, qt[:4], qt[4:])
    fac = fac[0] if isinstance(fac, tuple) else fac
    edge, plane = gold["assoc_edge"], gold["assoc_plane"]
    fe, fp = fac[fac["type"] == 1], fac[fac["type"] == 2]
    assert len(fe) == len(edge) and len(fp) == len(plane)
    assert np.array_equal(fe["p"], edge[:, 0:3]) and np.array_equal(fp["p"], plane[:, 0:3])
    assert np.abs(fe["a"] - edge[:, 3:6]).max() <= 1e-9 and np.abs(fe["b"] - edge[:, 6:9]).max() <= 1e-9
    assert np.abs(fp["a"] - plane[:, 3:6]).max() <= 1e-9 and np.abs(fp["b"][:, 0] - plane[:, 6]).max() <= 1e-9
    mc.close(), ms.close()


def test_eval_normal_eq_parity(ctx, oracle_mod, cfg_small):
    c = cfg_small
    mc, ms = _maps(ctx, c)
    qt = pose7(c["q0"], c["t0"])
    ctx.associate(mc, ms, c["corner"], c["surf"], c["q0"], c["t0"])
    want_f = oracle_mod.associate(c["map_corner"], c["map_surf"], c["corner"], c["surf"], qt)
    rng = np.random.default_rng(2)
    for trial in range(4):
        q = c["q0"] + rng.normal(0, 0.01, 4) * (trial > 0)
        q /= np.linalg.norm(q)
        t = c["t0"] + rng.normal(0, 0.1, 3) * (trial > 0)
        for huber in (0.1, 0.0):
            cost, H, g = ctx.eval_normal_eq(q, t, huber)
            wc, wH, wg = oracle_mod.evaluate(want_f, pose7(q, t), huber)
            assert abs(cost - wc) <= REL * abs(wc)
            assert np.abs(H - wH).max() <= REL * np.abs(wH).max()
            assert np.abs(g - wg).max() <= REL * np.abs(wg).max()
            # far inside the stated tolerance in practice
            assert abs(cost - wc) <= 1e-10 * abs(wc)
    mc.close(), ms.close()


def test_eval_normal_eq_bulk_kernel_parity(ctx, oracle_mod, cfg_small):
    """Above one 352-factor tile per SM the J^T J evaluation switches to the bulk-copy (TMA) staged kernel: same sums
    as the oracle (1e-5 relative stated, ~1e-12 observed), run-to-run identical, ragged last tile and a tile that
    straddles the corner / surf boundary included."""
    c = cfg_small
    mc, ms = _maps(ctx, c)
    rng = np.random.default_rng(5)
    corner = c["corner"][rng.integers(0, len(c["corner"]), 9001)]          # 25.6 tiles: the 26th is mixed
    surf = c["surf"][rng.integers(0, len(c["surf"]), 148 * 384 + 12345)]   # > one tile per SM, ragged tail
    qt = pose7(c["q0"], c["t0"])
    ctx.associate(mc, ms, corner, surf, c["q0"], c["t0"])
    want_f = oracle_mod.associate(c["map_corner"], c["map_surf"], corner, surf, qt)
    assert (want_f["type"] == 1).sum() > 1000 and (want_f["type"] == 2).sum() > 10000 and (want_f["type"] == 0).sum() > 100
    for huber in (0.1, 0.0):
        cost, H, g = ctx.eval_normal_eq(c["q0"], c["t0"], huber)
        wc, wH, wg = oracle_mod.evaluate(want_f, qt, huber)
        assert abs(cost - wc) <= 1e-10 * abs(wc)
        assert np.abs(H - wH).max() <= 1e-10 * np.abs(wH).max()
        assert np.abs(g - wg).max() <= 1e-9 * np.abs(wg).max()
        cost2, H2, g2 = ctx.eval_normal_eq(c["q0"], c["t0"], huber)
        assert cost2 == cost and np.array_equal(H2, H) and np.array_equal(g2, g)
    mc.close(), ms.close()


@pytest.mark.parametrize("max_iter", [0, 1, 4, 10, 50])
def test_solve_parity(ctx, oracle_mod, cfg_small, max_iter):
    """Device-resident Levenberg-Marquardt == restated Ceres loop: same accept/reject sequence, same termination."""
    c = cfg_small
    mc, ms = _maps(ctx, c)
    qt = pose7(c["q0"], c["t0"])
    ctx.associate(mc, ms, c["corner"], c["surf"], c["q0"], c["t0"])
    want_f = oracle_mod.associate(c["map_corner"], c["map_surf"], c["corner"], c["surf"], qt)
    q, t, s = ctx.solve(c["q0"], c["t0"], max_iter, 0.1)
    wx, ws = oracle_mod.solve(want_f, qt, max_iter, 0.1)
    assert s.termination == ws.termination
    assert s.iterations == ws.iterations
    assert s.num_successful_steps == ws.num_successful and s.num_unsuccessful_steps == ws.num_unsuccessful
    assert s.num_evaluations == ws.num_evals
    assert abs(s.initial_cost - ws.initial_cost) <= REL * ws.initial_cost
    assert abs(s.final_cost - ws.final_cost) <= REL * ws.final_cost
    assert np.linalg.norm(t - wx[4:]) < POSE_M
    assert np.abs(q - wx[:4]).max() < POSE_RAD / 2
    mc.close(), ms.close()


def _register_both(ctx, oracle_mod, ilsm, c, **kw):
    mc, ms = _maps(ctx, c)
    opts = ilsm.default_opts(**kw)
    q, t, rep = ctx.register(mc, ms, c["corner"], c["surf"], c["q0"], c["t0"], opts)
    wx, wsum, wnf = oracle_mod.register_aloam(c["map_corner"], c["map_surf"], c["corner"], c["surf"], pose7(c["q0"], c["t0"]),
                                             outer=opts.outer_iterations, max_iter=opts.max_num_iterations)
    mc.close(), ms.close()
    return q, t, rep, wx, wsum, wnf


def _check_registration(ilsm, q, t, rep, wx, wsum, wnf):
    assert rep.passes == len(wsum)
    for p in range(rep.passes):
        g, w = rep.pass_[p], wsum[p]
        assert g.num_edge_factors == wnf[2 * p] and g.num_plane_factors == wnf[2 * p + 1]
        assert g.termination == w.termination and g.iterations == w.iterations
        assert abs(g.final_cost - w.final_cost) <= REL * w.final_cost
    assert np.linalg.norm(t - wx[4:]) < POSE_M
    assert ilsm.synth.quat_angle(q, wx[:4]) < POSE_RAD


def test_register_parity_small(ctx, oracle_mod, ilsm, cfg_small):
    out = _register_both(ctx, oracle_mod, ilsm, cfg_small)
    _check_registration(ilsm, *out)
    q, t = out[0], out[1]
    assert np.linalg.norm(t - cfg_small["t_true"]) < 0.03  # and it actually registers the frame


def test_register_parity_config1_full(ctx, oracle_mod, ilsm, cfg_full):
    """BASELINE config 1 at full size: OS0-64 frame vs 100k-point map, laserMapping settings (2 x <=4 iterations)."""
    out = _register_both(ctx, oracle_mod, ilsm, cfg_full)
    _check_registration(ilsm, *out)


def test_register_mapoptimization_settings(ctx, oracle_mod, ilsm, cfg_small):
    """mapOptimization.cpp:377-450: plane factors only, one pass, max 10 iterations, termination type reported."""
    c = dict(cfg_small)
    c["corner"] = np.zeros((0, 4), np.float32)
    mc, ms = _maps(ctx, c)
    opts = ilsm.default_opts(outer_iterations=1, max_num_iterations=10, min_corner_map=0, min_surf_map=0)
    q, t, rep = ctx.register(mc, ms, c["corner"], c["surf"], c["q0"], c["t0"], opts)
    want_f = oracle_mod.associate(c["map_corner"], c["map_surf"], c["corner"], c["surf"], pose7(c["q0"], c["t0"]))
    wx, ws = oracle_mod.solve(want_f, pose7(c["q0"], c["t0"]), 10, 0.1)
    g = rep.pass_[0]
    assert g.num_edge_factors == 0 and g.num_plane_factors == (want_f["type"] == 2).sum()
    assert g.termination == ws.termination and g.iterations == ws.iterations
    assert np.linalg.norm(t - wx[4:]) < POSE_M and ilsm.synth.quat_angle(q, wx[:4]) < POSE_RAD
    mc.close(), ms.close()


def test_register_guard_and_empty_inputs(ctx, ilsm, cfg_small):
    c = cfg_small
    mc = ctx.new_map().set_input_cloud(c["map_corner"][:10])
    ms = ctx.new_map().set_input_cloud(c["map_surf"])
    with pytest.raises(ilsm.IlsmError) as e:  # laserMapping.cpp:624
        ctx.register(mc, ms, c["corner"], c["surf"], c["q0"], c["t0"])
    assert e.value.code == -4
    # no feature points at all: Ceres reports CONVERGENCE with the pose untouched
    mc.set_input_cloud(c["map_corner"])
    q, t, rep = ctx.register(mc, ms, np.zeros((0, 4), np.float32), np.zeros((0, 4), np.float32), c["q0"], c["t0"])
    assert np.array_equal(q, c["q0"]) and np.array_equal(t, c["t0"])
    assert rep.pass_[0].termination == ilsm.CONVERGENCE and rep.pass_[0].iterations == 0
    mc.close(), ms.close()


def test_register_is_deterministic(ctx, ilsm, cfg_small):
    c = cfg_small
    mc, ms = _maps(ctx, c)
    outs = [ctx.register(mc, ms, c["corner"], c["surf"], c["q0"], c["t0"]) for _ in range(3)]
    for q, t, _ in outs[1:]:
        assert np.array_equal(q, outs[0][0]) and np.array_equal(t, outs[0][1])
    mc.close(), ms.close()


def test_register_dev_matches_host_entry(ctx, ilsm, cfg_small):
    import torch
    c = cfg_small
    mc, ms = _maps(ctx, c)
    q, t, rep = ctx.register(mc, ms, c["corner"], c["surf"], c["q0"], c["t0"])
    dc = torch.from_numpy(c["corner"]).cuda()
    ds = torch.from_numpy(c["surf"]).cuda()
    pose = torch.from_numpy(pose7(c["q0"], c["t0"])).cuda()
    torch.cuda.synchronize()
    ctx.register_dev(mc, ms, dc.data_ptr(), len(dc), ds.data_ptr(), len(ds), 16, pose.data_ptr())
    ctx.sync()
    out = pose.cpu().numpy()
    assert np.array_equal(out[:4], q) and np.array_equal(out[4:], t)
    mc.close(), ms.close()


def test_knn_matches_reference_nanoflann_golden(ctx):
    """The CUDA k-NN against the committed outputs of the reference's own vendored nanoflann (tests/golden)."""
    import os
    g = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "knn_nanoflann.npz"))
    for cell in (0.0, 0.7):
        lm = ctx.new_map().set_input_cloud(g["map"], cell)
        for k in (1, 5, 8):
            idx, d2 = lm.nearest_k_search(g["queries"], k)
            assert np.array_equal(idx, g[f"idx_k{k}"]) and np.array_equal(d2, g[f"d2_k{k}"]), (cell, k)
        lm.close()


def test_register_matches_oracle_regression_golden(ctx, ilsm):
    import os
    g = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "oracle_regression.npz"))
    cs = ilsm.synth.config1(n_map=6000)
    mc, ms = ctx.new_map().set_input_cloud(cs["map_corner"]), ctx.new_map().set_input_cloud(cs["map_surf"])
    q, t, rep = ctx.register(mc, ms, cs["corner"], cs["surf"], cs["q0"], cs["t0"])
    assert np.linalg.norm(t - g["pose"][4:]) < POSE_M and ilsm.synth.quat_angle(q, g["pose"][:4]) < POSE_RAD
    assert [rep.pass_[p].termination for p in range(2)] == list(g["term"])
    assert [rep.pass_[p].num_edge_factors for p in range(2)] == [int(g["factors"][0]), int(g["factors"][2])]
    assert [rep.pass_[p].num_plane_factors for p in range(2)] == [int(g["factors"][1]), int(g["factors"][3])]
    f = ctx.extract_features(cs["cloud"])
    for key in ("sharp_idx", "less_sharp_idx", "flat_idx"):
        assert np.array_equal(f[key], g[key]), key
    assert len(f["less_flat"]) == int(g["n_less_flat"])
    mc.close(), ms.close()


def test_knn_config3_full_size_sample(ctx, oracle_mod, ilsm):
    """BASELINE configs[2] at its largest point: a 2M-point map, the whole 65536-ray frame as queries.  The oracle k-d tree
    answers a 768-query sample (bit-exact indices and distances), and two size-independent properties hold for all of
    them: ascending distances per query, and d2 recomputed from the returned index equals the returned d2."""
    S = ilsm.synth
    c = S.config1(n_map=2_000_000)
    m = np.ascontiguousarray(np.concatenate([c["map_corner"], c["map_surf"]])[:, :3], np.float32)
    R = S.quat_to_mat(c["q_true"])
    q = (c["cloud"][:, :3].astype(np.float64) @ R.T + c["t_true"]).astype(np.float32)
    lm = ctx.new_map().set_input_cloud(m)
    idx, d2 = lm.nearest_k_search(q, 5)
    assert (idx >= 0).all() and (np.diff(d2, axis=1) >= 0).all()
    nb = m[idx]                                               # (Q, 5, 3)
    dx, dy, dz = (q[:, None, 0] - nb[..., 0]), (q[:, None, 1] - nb[..., 1]), (q[:, None, 2] - nb[..., 2])
    assert np.array_equal((dx * dx + dy * dy) + dz * dz, d2)  # FLANN L2_Simple order, float, no FMA
    sel = np.random.default_rng(3).choice(len(q), 768, replace=False)
    ri, rd = oracle_mod.knn_kdtree(m, q[sel], 5)
    assert np.array_equal(idx[sel], ri) and np.array_equal(d2[sel], rd)
    lm.close()


def test_eval_normal_eq_is_additive_over_factor_sets(ctx, cfg_small):
    """Size-independent property of the J^T J kernels: the sums over k copies of a factor set are k times the sums over
    one copy (small kernel vs bulk-copy kernel, 1e-12 relative: different summation orders)."""
    c = cfg_small
    mc, ms = _maps(ctx, c)
    ctx.associate(mc, ms, c["corner"], c["surf"], c["q0"], c["t0"])
    cost1, H1, g1 = ctx.eval_normal_eq(c["q0"], c["t0"], 0.1)
    k = 1 + (148 * 352) // (len(c["corner"]) + len(c["surf"]))  # enough copies for one tile per SM: the bulk kernel
    ctx.associate(mc, ms, np.tile(c["corner"], (k, 1)), np.tile(c["surf"], (k, 1)), c["q0"], c["t0"])
    costk, Hk, gk = ctx.eval_normal_eq(c["q0"], c["t0"], 0.1)
    assert abs(costk - k * cost1) <= 1e-12 * k * cost1
    assert np.abs(Hk - k * H1).max() <= 1e-12 * k * np.abs(H1).max()
    assert np.abs(gk - k * g1).max() <= 1e-11 * k * np.abs(g1).max()
    mc.close(), ms.close()


def test_gpu_evaluation_matches_reference_functors(ctx, oracle_mod, cfg_small):
    """The CUDA evaluation (cost, J^T J, J^T r) against the REFERENCE'S OWN Ceres functors
    (lidarFeaturePointsFunction.hpp on dual numbers, oracle/_ref/libref_functors.so -- prebuilt, it travels with the
    snapshot): factors from the GPU association, the loss switched off (huber_a = 0), and with HuberLoss(0.1) applied
    to the reference residual blocks through Ceres' corrector (rho'' <= 0: residual and Jacobian scaled by sqrt(rho'))."""
    if oracle_mod.ref_functors() is None:
        pytest.skip("oracle/_ref/libref_functors.so not available")
    from test_ref_functors_cpu import tangent
    c = cfg_small
    qt = np.concatenate([c["q0"], c["t0"]])
    mc = ctx.new_map().set_input_cloud(c["map_corner"])
    ms = ctx.new_map().set_input_cloud(c["map_surf"])
    fac = ctx.associate(mc, ms, c["corner"], c["surf"], c["q0"], c["t0"])
    assert (fac["type"] == 1).sum() > 20 and (fac["type"] == 2).sum() > 100
    for huber in (0.0, 0.1):
        cost_r, H_r, g_r = 0.0, np.zeros((6, 6)), np.zeros(6)
        for rec in fac[fac["type"] != 0]:
            r, J = oracle_mod.ref_functor_eval(int(rec["type"]), rec["p"], rec["a"], rec["b"], (0, 0, 0), 1.0, qt)
            Jt = tangent(J, qt[:4])
            s = float(r @ r)
            rho0, sc = s, 1.0
            if huber > 0 and s > huber * huber:
                rho0, sc = 2 * huber * np.sqrt(s) - huber * huber, np.sqrt(huber / np.sqrt(s))
            cost_r += 0.5 * rho0
            H_r += (sc * Jt).T @ (sc * Jt)
            g_r += (sc * Jt).T @ (sc * r)
        cost, H, g = ctx.eval_normal_eq(c["q0"], c["t0"], huber)
        # north_star: residuals within 1e-5 relative (observed ~1e-13)
        assert abs(cost - cost_r) <= 1e-9 * cost_r, huber
        assert np.abs(H - H_r).max() <= 1e-9 * np.abs(H_r).max() and np.abs(g - g_r).max() <= 1e-9 * np.abs(g_r).max(), huber
    mc.close(), ms.close()


@pytest.mark.parametrize("k,max_dist", [(5, 0.0), (1, 0.0), (8, 0.0), (5, 1.0)])
def test_knn_binned_path_equals_per_query_path(ilsm, ctx, oracle_mod, cfg_small, k, max_dist):
    """The query-binned search (one warp per group of <= 32 queries of a voxel, knn_binned.cu) forced onto a mixed query
    set -- groups of every size from 1 to 70 queries per voxel, exact duplicates, far-away queries that need ring
    expansion or the brute-force sweep, NaN / out-of-range queries -- must return exactly what the per-query kernel and
    the oracle return."""
    import os
    rng = np.random.default_rng(1000 + k)
    m = cfg_small["map_surf"]
    parts = []
    anchors = m[rng.integers(0, len(m), 120)]
    for i, a in enumerate(anchors):                       # i + 1 queries inside one 1 m voxel around a map point
        base = np.floor(a[:3]) + 0.5
        parts.append(base + rng.uniform(-0.49, 0.49, ((i % 70) + 1, 3)))
    parts.append(np.repeat(m[5:6, :3], 40, axis=0))       # 40 identical queries
    parts.append(rng.uniform(-300, 300, (200, 3)))        # mostly empty neighbourhoods
    parts.append(np.array([[np.nan, 0, 0], [0, np.inf, 0], [3e6, 0, 0], [0, 0, -2e6]]))
    q = np.concatenate(parts).astype(np.float32)
    rng.shuffle(q)
    os.environ["ILSM_KNN_BINNED_MIN"] = "1"
    try:
        bctx = ilsm.Context(0)
    finally:
        del os.environ["ILSM_KNN_BINNED_MIN"]
    bm = bctx.new_map().set_input_cloud(m)
    pm = ctx.new_map().set_input_cloud(m)
    bi, bd = bm.nearest_k_search(q, k, max_dist=max_dist)
    pi, pd = pm.nearest_k_search(q, k, max_dist=max_dist)
    finite = np.isfinite(q).all(axis=1)
    if max_dist > 0:
        ins = pd < max_dist * max_dist
        assert np.array_equal(bi[ins], pi[ins]) and np.array_equal(bd[ins], pd[ins])
    else:
        assert np.array_equal(bi[finite], pi[finite]) and np.array_equal(bd[finite], pd[finite])
        ok = finite & (np.abs(q) < 1e6).all(axis=1)
        ri, rd = oracle_mod.knn_kdtree(m, q[ok], k)
        assert np.array_equal(bi[ok], ri) and np.array_equal(bd[ok], rd)
    # a map with fewer points than k, and a rebuilt query binning on the same context
    tiny = bctx.new_map().set_input_cloud(m[:3])
    ti, td = tiny.nearest_k_search(q[:50], k)
    kk = min(k, 3)
    f50 = np.isfinite(q[:50]).all(axis=1)
    assert (ti[f50][:, kk:] == -1).all() and (ti[f50][:, :kk] >= 0).all()
    bm.close(), pm.close(), tiny.close(), bctx.close()


def test_map_rebuilds_alternate_tables_and_pair_build(ctx, oracle_mod, ilsm, cfg_small):
    """The map alternates between two hash tables (each build's scatter launch empties the other one): a sequence of
    rebuilds with growing, shrinking, empty and regrown clouds on ONE handle must answer every query set exactly, and
    the pair build (both structures of a frame in one set of launches) must equal two single builds."""
    import torch
    rng = np.random.default_rng(77)
    base = cfg_small["map_surf"][:, :3]
    m = ctx.new_map()
    for n in (3000, 17000, 500, 0, 9000, 9000, 40, 17000, 2500):
        pts = (base[rng.choice(len(base), n, replace=False)] + rng.normal(0, 0.02, (n, 3))).astype(np.float32) if n else np.zeros((0, 3), np.float32)
        m.set_input_cloud(pts)
        q = (base[rng.integers(0, len(base), 300)] + rng.normal(0, 0.4, (300, 3))).astype(np.float32)
        idx, d2 = m.nearest_k_search(q, 5)
        if n == 0:
            assert (idx == -1).all()
            continue
        ri, rd = oracle_mod.knn_kdtree(pts, q, 5)
        assert np.array_equal(idx, ri) and np.array_equal(d2, rd), n
    m.close()
    # pair build on device-resident clouds, repeated (the tables alternate), against single builds
    dev = torch.device("cuda:0")
    a, b = ctx.new_map(), ctx.new_map()
    sa, sb = ctx.new_map(), ctx.new_map()
    for rep, (na, nb) in enumerate(((4000, 12000), (12000, 300), (0, 5000), (7000, 0), (6000, 6000))):
        pa = np.zeros((na, 4), np.float32)
        pb = np.zeros((nb, 4), np.float32)
        pa[:, :3] = cfg_small["map_corner"][rng.choice(len(cfg_small["map_corner"]), na, replace=na > len(cfg_small["map_corner"]))][:, :3]
        pb[:, :3] = base[rng.choice(len(base), nb, replace=False)]
        da, db = torch.from_numpy(pa).to(dev), torch.from_numpy(pb).to(dev)
        a.build_pair_dev(da.data_ptr(), na, b, db.data_ptr(), nb, 16)
        ctx.sync()
        sa.set_input_cloud(pa), sb.set_input_cloud(pb)
        q = (base[rng.integers(0, len(base), 400)] + rng.normal(0, 0.5, (400, 3))).astype(np.float32)
        for pm, sm_ in ((a, sa), (b, sb)):
            i1, d1 = pm.nearest_k_search(q, 5)
            i2, d2_ = sm_.nearest_k_search(q, 5)
            assert np.array_equal(i1, i2) and np.array_equal(d1, d2_), rep
    for x in (a, b, sa, sb):
        x.close()
