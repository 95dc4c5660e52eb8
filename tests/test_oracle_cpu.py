"""CPU suite: pins the oracle (k-NN against the reference's own vendored nanoflann; fits / Jacobians / LM against
independent numpy / scipy / torch-autograd restatements).  No GPU needed."""
import numpy as np
import pytest

from conftest import pose7


def test_knn_matches_reference_nanoflann(oracle_mod, cfg_small):
    """oracle brute force == oracle k-d tree == nanoflann 1.3.2 from /root/reference/include (oracle/_ref)."""
    o = oracle_mod
    if o.ref() is None:
        pytest.skip("oracle/_ref not built (reference tree absent and no prebuilt library)")
    rng = np.random.default_rng(11)
    m = cfg_small["map_surf"]
    q = m[rng.integers(0, len(m), 1500)] + rng.normal(0, 0.4, (1500, 3)).astype(np.float32)
    ib, db = o.knn_brute(m, q, 5)
    ik, dk = o.knn_kdtree(m, q, 5)
    ir, dr = o.RefKdTree(m, leaf_max=10).knn(q, 5)
    assert np.array_equal(ib, ik) and np.array_equal(db, dk)
    assert np.array_equal(db, dr)
    assert np.array_equal(ib, ir)
    i1, d1 = o.knn_kdtree(m, q, 1)
    assert np.array_equal(i1[:, 0], ib[:, 0]) and np.array_equal(d1[:, 0], db[:, 0])


def test_knn_tie_break_lower_index(oracle_mod):
    o = oracle_mod
    m = np.array([[1, 0, 0], [0, 1, 0], [-1, 0, 0], [0, -1, 0], [0, 0, 1], [0, 0, -1], [5, 5, 5]], np.float32)
    q = np.zeros((1, 3), np.float32)
    for fn in (o.knn_brute, o.knn_kdtree):
        idx, d2 = fn(m, q, 5)
        assert idx[0].tolist() == [0, 1, 2, 3, 4] and np.all(d2[0] == 1.0)


def test_knn_ragged(oracle_mod):
    o = oracle_mod
    m = np.array([[0, 0, 0], [1, 0, 0]], np.float32)
    idx, d2 = o.knn_kdtree(m, np.array([[0.2, 0, 0]], np.float32), 5)
    assert idx[0].tolist() == [0, 1, -1, -1, -1] and np.isinf(d2[0, 2:]).all()


def test_transform_is_double_math_float_store(oracle_mod):
    o = oracle_mod
    rng = np.random.default_rng(3)
    q = rng.normal(size=4)
    q /= np.linalg.norm(q)
    t = rng.normal(size=3) * 10
    p = rng.normal(size=(200, 3)).astype(np.float32) * 30
    x, y, z, w = q
    R = np.array([[1 - 2 * (y * y + z * z), 2 * (x * y - z * w), 2 * (x * z + y * w)],
                  [2 * (x * y + z * w), 1 - 2 * (x * x + z * z), 2 * (y * z - x * w)],
                  [2 * (x * z - y * w), 2 * (y * z + x * w), 1 - 2 * (x * x + y * y)]])
    want = (p.astype(np.float64) @ R.T + t)
    got = o.transform_points(pose7(q, t), p)
    assert got.dtype == np.float32
    assert np.max(np.abs(got - want)) < 1e-5


def test_line_fit_vs_numpy_eigh(oracle_mod):
    o = oracle_mod
    rng = np.random.default_rng(5)
    n_ok = 0
    for _ in range(300):
        c = rng.uniform(-100, 100, 3)
        d = rng.normal(size=3)
        d /= np.linalg.norm(d)
        nb = (c + np.outer(rng.uniform(-1, 1, 5), d) + rng.normal(0, rng.choice([0.01, 0.3]), (5, 3))).astype(np.float32)
        ok, f = o.fit_line(nb)
        P = nb.astype(np.float64)
        cen = P.sum(0) / 5.0
        w, V = np.linalg.eigh((P - cen).T @ (P - cen))
        assert ok == bool(w[2] > 3 * w[1])
        if ok:
            n_ok += 1
            v = V[:, 2]
            ab = np.stack([f["a"], f["b"]])
            want = np.stack([cen + 0.1 * v, cen - 0.1 * v])
            err = min(np.abs(ab - want).max(), np.abs(ab - want[::-1]).max())
            assert err < 1e-9
    assert n_ok > 50


def test_plane_fit_vs_numpy_lstsq(oracle_mod):
    o = oracle_mod
    rng = np.random.default_rng(6)
    n_ok = 0
    for _ in range(300):
        n = rng.normal(size=3)
        n /= np.linalg.norm(n)
        c = rng.uniform(-100, 100, 3)
        u = np.cross(n, rng.normal(size=3))
        u /= np.linalg.norm(u)
        v = np.cross(n, u)
        ab = rng.uniform(-0.8, 0.8, (5, 2))
        nb = (c + ab[:, :1] * u + ab[:, 1:] * v + rng.normal(0, rng.choice([0.01, 0.15]), (5, 1)) * n).astype(np.float32)
        ok, f = o.fit_plane(nb)
        A = nb.astype(np.float64)
        x = np.linalg.lstsq(A, -np.ones(5), rcond=None)[0]
        nn = np.linalg.norm(x)
        want_ok = bool(np.all(np.abs(A @ (x / nn) + 1 / nn) <= 0.2))
        assert ok == want_ok
        if ok:
            n_ok += 1
            assert np.abs(f["a"] - x / nn).max() < 1e-8
            assert abs(f["b"][0] - 1 / nn) < 1e-7 * max(1.0, 1 / nn)
    assert n_ok > 50


def _functor_residuals_torch(f, qt):
    """lidarFeaturePointsFunction.hpp:199-293 written with torch in float64 on the AMBIENT parameters (q xyzw, t)."""
    import torch
    q, t = qt[:4], qt[4:]
    u, w = q[:3], q[3]
    p = torch.tensor(f["p"], dtype=torch.float64)
    uv = 2 * torch.linalg.cross(u, p)
    lp = p + w * uv + torch.linalg.cross(u, uv) + t
    a = torch.tensor(f["a"], dtype=torch.float64)
    b = torch.tensor(f["b"], dtype=torch.float64)
    if f["type"] == 1:
        nu = torch.linalg.cross(lp - a, lp - b)
        return nu / torch.linalg.norm(a - b)
    return (a @ lp + b[0]).reshape(1)


def test_jacobians_vs_autograd(oracle_mod, cfg_small):
    """Analytic tangent Jacobians == autograd through the functor formulas times the
    EigenQuaternionParameterization plus-Jacobian."""
    import torch
    o = oracle_mod
    c = cfg_small
    qt = pose7(c["q0"], c["t0"])
    fac = o.associate(c["map_corner"], c["map_surf"], c["corner"], c["surf"], qt)
    edges = fac[fac["type"] == 1][:6]
    planes = fac[fac["type"] == 2][:6]
    assert len(edges) and len(planes)
    x, y, z, w = qt[:4]
    plusJ = np.array([[w, z, -y], [-z, w, x], [y, -x, w], [-x, -y, -z]])
    for f in list(edges) + list(planes):
        xt = torch.tensor(qt, dtype=torch.float64, requires_grad=True)
        J = torch.autograd.functional.jacobian(lambda v: _functor_residuals_torch(f, v), xt).numpy()
        Jl = np.concatenate([J[:, :4] @ plusJ, J[:, 4:]], axis=1)
        r = _functor_residuals_torch(f, xt).detach().numpy()
        one = np.array([f], dtype=fac.dtype)
        cost, H, g, res = o.evaluate(one, qt, huber_a=0.0, want_residuals=True)
        assert np.allclose(res[0, :len(r)], r, rtol=1e-12, atol=1e-12)
        assert np.allclose(H, Jl.T @ Jl, rtol=1e-9, atol=1e-9)
        assert np.allclose(g, Jl.T @ r, rtol=1e-9, atol=1e-9)
        assert np.isclose(cost, 0.5 * r @ r, rtol=1e-12)


def test_huber_corrector(oracle_mod, cfg_small):
    o = oracle_mod
    c = cfg_small
    qt = pose7(c["q0"], c["t0"])
    fac = o.associate(c["map_corner"], c["map_surf"], c["corner"], c["surf"], qt)
    fac = fac[fac["type"] != 0]
    cost, H, g, res = o.evaluate(fac, qt, huber_a=0.1, want_residuals=True)
    s = (res ** 2).sum(1)
    rho = np.where(s > 0.01, 2 * 0.1 * np.sqrt(s) - 0.01, s)
    assert np.isclose(cost, 0.5 * rho.sum(), rtol=1e-12)
    assert (s > 0.01).any() and (s <= 0.01).any()


def test_lm_fixed_point_vs_scipy(oracle_mod, cfg_small):
    """With the factor set frozen, the restated Ceres LM run to convergence lands on the same minimiser as
    scipy.optimize.least_squares(loss='huber') on the same residuals (independent solver)."""
    from scipy.optimize import least_squares
    o = oracle_mod
    c = cfg_small
    qt0 = pose7(c["q0"], c["t0"])
    fac = o.associate(c["map_corner"], c["map_surf"], c["corner"], c["surf"], qt0)
    fac = fac[fac["type"] != 0]
    x, s = o.solve(fac, qt0, max_iter=50)
    assert s.termination == 0 and s.final_cost < s.initial_cost

    def plus(qt, d):
        nd = np.linalg.norm(d[:3])
        dq = np.array([0, 0, 0, 1.0]) if nd == 0 else np.concatenate([np.sin(nd) / nd * d[:3], [np.cos(nd)]])
        ax, ay, az, aw = dq
        bx, by, bz, bw = qt[:4]
        q = np.array([aw * bx + ax * bw + ay * bz - az * by, aw * by + ay * bw + az * bx - ax * bz,
                      aw * bz + az * bw + ax * by - ay * bx, aw * bw - ax * bx - ay * by - az * bz])
        return np.concatenate([q, qt[4:] + d[3:]])

    def blocks(d):
        _, _, _, res = o.evaluate(fac, plus(x, d), huber_a=0.0, want_residuals=True)
        return res

    # scipy's huber acts per scalar residual; Ceres' acts per residual block.  Minimise the block-Huber cost by
    # feeding scipy sqrt(rho(|r_block|^2)) as a scalar residual per block with a linear loss.
    def fun(d):
        s2 = (blocks(d) ** 2).sum(1)
        rho = np.where(s2 > 0.01, 2 * 0.1 * np.sqrt(s2) - 0.01, s2)
        return np.sqrt(rho)

    sol = least_squares(fun, np.zeros(6), method="trf", xtol=1e-14, ftol=1e-14, gtol=1e-14)
    # Ceres stops on |d cost| <= 1e-6 * cost, so the oracle sits within that band of the true minimum
    best = 0.5 * (fun(sol.x) ** 2).sum()
    assert best <= s.final_cost * (1 + 1e-12)
    assert s.final_cost - best <= 2e-6 * s.final_cost
    assert np.linalg.norm(sol.x[3:]) < 5e-4 and np.linalg.norm(sol.x[:3]) < 5e-4


def test_register_converges_to_truth(oracle_mod, ilsm, cfg_small):
    o = oracle_mod
    c = cfg_small
    x, sums, nf = o.register_aloam(c["map_corner"], c["map_surf"], c["corner"], c["surf"], pose7(c["q0"], c["t0"]))
    assert len(sums) == 2 and nf[0] > 50 and nf[1] > 200
    assert np.linalg.norm(x[4:] - c["t_true"]) < 0.03
    assert ilsm.synth.quat_angle(x[:4], c["q_true"]) < 2e-3
    assert sums[0].iterations <= 4 and sums[0].num_evals <= 5


def test_register_guard_and_empty(oracle_mod, cfg_small):
    o = oracle_mod
    c = cfg_small
    qt = pose7(c["q0"], c["t0"])
    x, sums, nf = o.register_aloam(c["map_corner"][:10], c["map_surf"], c["corner"], c["surf"], qt)
    assert len(sums) == 0 and np.array_equal(x, qt)  # laserMapping.cpp:624 guard
    x, s = o.solve(np.zeros(0, o.FACTOR_DTYPE), qt)
    assert s.termination == 0 and s.iterations == 0 and np.array_equal(x, qt)
