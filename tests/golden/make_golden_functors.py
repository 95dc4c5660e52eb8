#!/usr/bin/env python
"""Generates tests/golden/functors_reference.npz from the REFERENCE'S OWN Ceres functors (run in the build container):

  python tests/golden/make_golden_functors.py

/root/reference/src/lidarFeaturePointsFunction.hpp is compiled from where it lies into oracle/_ref/libref_functors.so
(oracle/Makefile; Eigen / Ceres stand-ins in oracle/shims/, dual numbers in oracle/ref_functors.cpp) and evaluated on
random factors of the four functor types the path uses, at random poses: residuals and the ambient Jacobian
d r / d (qx, qy, qz, qw, tx, ty, tz) that ceres::AutoDiffCostFunction would hand to the solver.
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)

import oracle  # noqa: E402


def main():
    assert oracle.ref_functors() is not None, "oracle/_ref/libref_functors.so not built (needs /root/reference)"
    rng = np.random.default_rng(0xF0C7)
    n = 240
    ftype = np.repeat([1, 2, 3, 4], n // 4).astype(np.int32)
    p = rng.uniform(-40, 40, (n, 3))
    a = p + rng.normal(0, 1.0, (n, 3))
    b = a + rng.normal(0, 0.3, (n, 3))
    c = a + rng.normal(0, 0.3, (n, 3))
    nrm = rng.normal(0, 1, (n, 3))
    nrm /= np.linalg.norm(nrm, axis=1, keepdims=True)
    pl = ftype == 2
    a[pl] = nrm[pl]
    b[pl, 0] = rng.uniform(-5, 5, pl.sum())
    qt = np.zeros((n, 7))
    ax = rng.normal(0, 1, (n, 3))
    ax /= np.linalg.norm(ax, axis=1, keepdims=True)
    ang = rng.uniform(-1.5, 1.5, n)
    ang[::7] = 0.0                      # identity rotations: the epsilon branch of Eigen's slerp
    qt[:, :3] = ax * np.sin(ang / 2)[:, None]
    qt[:, 3] = np.cos(ang / 2)
    qt[5::11, :4] *= -1.0               # w < 0: slerp's sign correction
    qt[:, 4:] = rng.uniform(-3, 3, (n, 3))
    r = np.zeros((n, 3))
    J = np.zeros((n, 3, 7))
    rows = np.zeros(n, np.int32)
    for i in range(n):
        ri, Ji = oracle.ref_functor_eval(ftype[i], p[i], a[i], b[i], c[i], 1.0, qt[i])
        rows[i] = len(ri)
        r[i, :len(ri)], J[i, :len(ri)] = ri, Ji
    assert np.isfinite(r).all() and np.isfinite(J).all()
    path = os.path.join(HERE, "functors_reference.npz")
    np.savez_compressed(path, ftype=ftype, p=p, a=a, b=b, c=c, qt=qt, rows=rows, r=r, J=J)
    print("functors_reference.npz", os.path.getsize(path), "bytes")


if __name__ == "__main__":
    main()
