"""SURVEY section 8 rows b4, b5, c6 (and f3's front_end_residual) pinned against the REFERENCE'S OWN Ceres functors.

tests/golden/functors_reference.npz holds residuals and ambient Jacobians of LidarEdgeFactor, LidarPlaneNormFactor,
front_end_residual and LidarPlaneFactor as written in /root/reference/src/lidarFeaturePointsFunction.hpp, evaluated on
dual numbers (oracle/ref_functors.cpp, generator tests/golden/make_golden_functors.py).  The oracle evaluates the same
factors with closed-form residuals and tangent-space Jacobians; composing the reference's ambient Jacobian with
EigenQuaternionParameterization's plus-Jacobian (Ceres 1.14, restated below: Plus(x, d) = [sin|d|/|d| d, cos|d|] * x)
must give the oracle's cost, J^T J and J^T r.  Ceres itself (the loss corrector, the solver) stays unpinned."""
import os

import numpy as np
import pytest

G = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "functors_reference.npz")


def plus_jacobian(q):
    """d Plus(x, delta) / d delta at delta = 0 for x = (x, y, z, w): column k = (e_k, 0) (x) x (Hamilton product)."""
    x, y, z, w = q
    return np.array([[w, z, -y], [-z, w, x], [y, -x, w], [-x, -y, -z]], float)


def tangent(J_amb, q):
    return np.concatenate([J_amb[:, :4] @ plus_jacobian(q), J_amb[:, 4:]], axis=1)


def to_oracle_factor(oracle_mod, ftype, p, a, b, c):
    f = np.zeros(1, oracle_mod.FACTOR_DTYPE)
    f["p"][0] = p
    if ftype == 1:
        f["type"], f["a"][0], f["b"][0] = 1, a, b
    elif ftype == 2:
        f["type"], f["a"][0], f["b"][0] = 2, a, [b[0], 0, 0]
    elif ftype == 3:
        f["type"], f["a"][0] = 3, a
    else:  # LidarPlaneFactor -> unit normal + offset, the form the association kernels emit (r = n.lp - n.j)
        n = np.cross(a - b, a - c)
        n = n / np.linalg.norm(n)
        f["type"], f["a"][0], f["b"][0] = 2, n, [-float(a @ n), 0, 0]
    return f


def check(oracle_mod, ftype, p, a, b, c, qt, r_ref, J_ref):
    Jt = tangent(J_ref, qt[:4])
    f = to_oracle_factor(oracle_mod, ftype, p, a, b, c)
    cost, H, g, res = oracle_mod.evaluate(f, qt, 0.0, want_residuals=True)
    nr = len(r_ref)
    scale = 1.0 + np.abs(r_ref).max()
    assert np.abs(res[0, :nr] - r_ref).max() <= 1e-9 * scale, (ftype, res[0], r_ref)
    assert abs(cost - 0.5 * float(r_ref @ r_ref)) <= 1e-9 * (1.0 + cost)
    Hr, gr = Jt.T @ Jt, Jt.T @ r_ref
    assert np.abs(H - Hr).max() <= 1e-9 * (1.0 + np.abs(Hr).max()), ftype
    assert np.abs(g - gr).max() <= 1e-9 * (1.0 + np.abs(gr).max()), ftype


def test_oracle_residuals_and_jacobians_match_reference_functors_golden(oracle_mod):
    g = np.load(G)
    assert set(g["ftype"].tolist()) == {1, 2, 3, 4}
    for i in range(len(g["ftype"])):
        n = int(g["rows"][i])
        check(oracle_mod, int(g["ftype"][i]), g["p"][i], g["a"][i], g["b"][i], g["c"][i], g["qt"][i], g["r"][i, :n], g["J"][i, :n])


def test_plus_jacobian_is_the_derivative_of_plus():
    """The restated EigenQuaternionParameterization Jacobian against a finite difference of Plus itself."""
    rng = np.random.default_rng(3)
    q = rng.normal(0, 1, 4)
    q /= np.linalg.norm(q)

    def qmul(a, b):
        return np.array([a[3] * b[0] + a[0] * b[3] + a[1] * b[2] - a[2] * b[1], a[3] * b[1] - a[0] * b[2] + a[1] * b[3] + a[2] * b[0],
                         a[3] * b[2] + a[0] * b[1] - a[1] * b[0] + a[2] * b[3], a[3] * b[3] - a[0] * b[0] - a[1] * b[1] - a[2] * b[2]])

    def plus(x, d):
        n = np.linalg.norm(d)
        dq = np.r_[np.sin(n) / n * d, np.cos(n)] if n > 0 else np.array([0, 0, 0, 1.0])
        return qmul(dq, x)

    P = plus_jacobian(q)
    for k in range(3):
        e = np.zeros(3)
        e[k] = 1e-6
        assert np.abs((plus(q, e) - plus(q, -e)) / 2e-6 - P[:, k]).max() < 1e-9


def test_oracle_matches_live_reference_functors_on_associated_factors(oracle_mod, cfg_small):
    """Factors produced by the oracle's own association of a config-1 frame (real line / plane fits), evaluated by the
    reference functors at the initial guess: the whole problem's cost, J^T J and J^T r without the loss."""
    if oracle_mod.ref_functors() is None:
        pytest.skip("oracle/_ref/libref_functors.so not available (built where /root/reference exists)")
    c = cfg_small
    qt = np.concatenate([c["q0"], c["t0"]])
    f = oracle_mod.associate(c["map_corner"], c["map_surf"], c["corner"], c["surf"], qt)
    f = f[f["type"] != 0][:400]
    assert (f["type"] == 1).sum() > 20 and (f["type"] == 2).sum() > 100
    cost_r, H_r, g_r = 0.0, np.zeros((6, 6)), np.zeros(6)
    for rec in f:
        r, J = oracle_mod.ref_functor_eval(int(rec["type"]), rec["p"], rec["a"], rec["b"], (0, 0, 0), 1.0, qt)
        Jt = tangent(J, qt[:4])
        cost_r += 0.5 * float(r @ r)
        H_r += Jt.T @ Jt
        g_r += Jt.T @ r
    cost, H, g = oracle_mod.evaluate(f, qt, 0.0)
    assert abs(cost - cost_r) <= 1e-10 * cost_r
    assert np.abs(H - H_r).max() <= 1e-10 * np.abs(H_r).max() and np.abs(g - g_r).max() <= 1e-10 * np.abs(g_r).max()
