"""CPU oracle bindings (TEST INFRASTRUCTURE -- never imported by the product path).

ctypes wrappers over ``oracle/_build/libilsm_oracle.so`` (the CPU restatement, ``ilsm_oracle.cpp``) and
``oracle/_ref/libref_nanoflann.so`` (the reference's own vendored nanoflann 1.3.2, the only reference
component that compiles here).  Only ``tests/``, ``__graft_entry__.smoke()`` and the ``cpu_baseline`` /
``--impl reference`` legs of ``bench.py`` may import this package.

Parity status: unpinned except k-NN (see header of ``ilsm_oracle.cpp``).
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = os.path.join(_HERE, "_build", "libilsm_oracle.so")
_REF = os.path.join(_HERE, "_ref", "libref_nanoflann.so")
_REF_IKD = os.path.join(_HERE, "_ref", "libref_ikd.so")
_REF_SCANREG = os.path.join(_HERE, "_ref", "libref_scanreg.so")
_REF_LASERODOM = os.path.join(_HERE, "_ref", "libref_laserodom.so")
_REF_LASERMAPPING = os.path.join(_HERE, "_ref", "libref_lasermapping.so")
_REF_SCANCONTEXT = os.path.join(_HERE, "_ref", "libref_scancontext.so")
_REF_IMAGEHANDLER = os.path.join(_HERE, "_ref", "libref_imagehandler.so")
_REF_FUN = os.path.join(_HERE, "_ref", "libref_functors.so")
_REF_ALOAM = os.path.join(_HERE, "_ref", "libref_aloam.so")


def build(force: bool = False) -> None:
    """Compile the oracle (and oracle/_ref when /root/reference is present)."""
    srcs = [os.path.join(_HERE, f) for f in os.listdir(_HERE) if f.endswith((".cpp", ".hpp"))]
    stale = (not os.path.exists(_LIB)) or any(os.path.getmtime(s) > os.path.getmtime(_LIB) for s in srcs)
    need_ref = ((not os.path.exists(_REF)) and os.path.exists("/root/reference/include/nanoflann.hpp")) or \
        ((not os.path.exists(_REF_IKD)) and os.path.exists("/root/reference/src/ikd-Tree/ikd_Tree.cpp")) or \
        ((not os.path.exists(_REF_FUN)) and os.path.exists("/root/reference/src/lidarFeaturePointsFunction.hpp")) or \
        ((not os.path.exists(_REF_ALOAM)) and os.path.exists("/root/reference/include/nanoflann.hpp"))
    if force or stale or need_ref:
        subprocess.run(["make", "-C", _HERE] + (["-B"] if force else []), check=True, capture_output=True)


class Factor(C.Structure):
    _fields_ = [("type", C.c_int32), ("src", C.c_int32), ("p", C.c_double * 3), ("a", C.c_double * 3),
                ("b", C.c_double * 3)]


FACTOR_DTYPE = np.dtype([("type", "<i4"), ("src", "<i4"), ("p", "<f8", 3), ("a", "<f8", 3), ("b", "<f8", 3)])
assert FACTOR_DTYPE.itemsize == C.sizeof(Factor) == 80


class SolveSummary(C.Structure):
    _fields_ = [("termination", C.c_int32), ("iterations", C.c_int32), ("num_successful", C.c_int32),
                ("num_unsuccessful", C.c_int32), ("initial_cost", C.c_double), ("final_cost", C.c_double),
                ("num_evals", C.c_int32), ("pad", C.c_int32)]


_lib = None
_ref = None


def lib():
    global _lib
    if _lib is None:
        build()
        _lib = C.CDLL(_LIB)
    return _lib


def ref():
    """The reference's vendored nanoflann (None when oracle/_ref was never built)."""
    global _ref
    if _ref is None:
        if not os.path.exists(_REF):
            build()
        if not os.path.exists(_REF):
            return None
        _ref = C.CDLL(_REF)
        _ref.ref_kdtree_build.restype = C.c_void_p
        _ref.ref_kdtree_build.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_int]
        _ref.ref_kdtree_free.argtypes = [C.c_void_p]
        _ref.ref_kdtree_knn.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_void_p]
    return _ref


def _f32(a):
    a = np.ascontiguousarray(a, dtype=np.float32)
    assert a.ndim == 2 and a.shape[1] >= 3
    return a


def _p(a):
    return a.ctypes.data_as(C.c_void_p)


def _stride(a):
    return a.strides[0] if len(a) else 4 * a.shape[1]


def knn_brute(map_xyz, q_xyz, k=5):
    m, q = _f32(map_xyz), _f32(q_xyz)
    idx = np.empty((len(q), k), np.int32)
    d2 = np.empty((len(q), k), np.float32)
    lib().orc_knn_brute(_p(m), len(m), m.strides[0], _p(q), len(q), q.strides[0], k, _p(idx), _p(d2))
    return idx, d2


def knn_kdtree(map_xyz, q_xyz, k=5):
    m, q = _f32(map_xyz), _f32(q_xyz)
    idx = np.empty((len(q), k), np.int32)
    d2 = np.empty((len(q), k), np.float32)
    lib().orc_knn_kdtree(_p(m), len(m), m.strides[0], _p(q), len(q), q.strides[0], k, _p(idx), _p(d2))
    return idx, d2


def knn_kdtree_omp(map_xyz, q_xyz, k=5, threads=0, tree="own"):
    """k-NN only, OpenMP over the host cores: (idx, d2, threads used, build seconds, query seconds)."""
    m, q = _f32(map_xyz), _f32(q_xyz)
    idx = np.empty((len(q), k), np.int32)
    d2 = np.empty((len(q), k), np.float32)
    L = lib() if tree == "own" else lib_nanoflann()
    if L is None:
        raise RuntimeError("oracle/_ref/libref_aloam.so not available")
    bs, qs = C.c_double(0), C.c_double(0)
    used = L.orc_knn_kdtree_omp(_p(m), len(m), m.strides[0], _p(q), len(q), q.strides[0], k, _p(idx), _p(d2), int(threads),
                                C.byref(bs), C.byref(qs))
    return idx, d2, used, bs.value, qs.value


class RefKdTree:
    """nanoflann::KDTreeSingleIndexAdaptor<L2_Simple<float>,3> from /root/reference/include (prebuilt)."""

    def __init__(self, map_xyz, leaf_max=10):
        self._m = _f32(map_xyz)
        self._r = ref()
        if self._r is None:
            raise RuntimeError("oracle/_ref/libref_nanoflann.so not built")
        self._h = self._r.ref_kdtree_build(_p(self._m), len(self._m), self._m.strides[0], leaf_max)

    def knn(self, q_xyz, k=5):
        q = _f32(q_xyz)
        idx = np.empty((len(q), k), np.int32)
        d2 = np.empty((len(q), k), np.float32)
        self._r.ref_kdtree_knn(self._h, _p(q), len(q), q.strides[0], k, _p(idx), _p(d2))
        return idx, d2

    def __del__(self):
        try:
            self._r.ref_kdtree_free(self._h)
        except Exception:
            pass


def transform_points(qt, pts):
    p = _f32(pts)
    qt = np.ascontiguousarray(qt, np.float64)
    out = np.empty((len(p), 3), np.float32)
    lib().orc_transform_points(_p(qt), _p(p), len(p), p.strides[0], _p(out))
    return out


def fit_line(nb5x3):
    nb = np.ascontiguousarray(nb5x3, np.float32).reshape(15)
    f = np.zeros(1, FACTOR_DTYPE)
    ok = lib().orc_fit_line(_p(nb), _p(f))
    return bool(ok), f[0]


def fit_plane(nb5x3):
    nb = np.ascontiguousarray(nb5x3, np.float32).reshape(15)
    f = np.zeros(1, FACTOR_DTYPE)
    ok = lib().orc_fit_plane(_p(nb), _p(f))
    return bool(ok), f[0]


def associate(map_corner, map_surf, corner, surf, qt):
    mc, ms, c, s = _f32(map_corner), _f32(map_surf), _f32(corner), _f32(surf)
    assert _stride(mc) == _stride(ms) and (len(c) == 0 or len(s) == 0 or _stride(c) == _stride(s))
    qt = np.ascontiguousarray(qt, np.float64)
    out = np.zeros(len(c) + len(s), FACTOR_DTYPE)
    n = lib().orc_associate(_p(mc), len(mc), _p(ms), len(ms), _stride(mc), _p(c), len(c), _p(s), len(s),
                            _stride(c) if len(c) else _stride(s), _p(qt), _p(out))
    return out[:n]


def evaluate(factors, qt, huber_a=0.1, want_residuals=False):
    f = np.ascontiguousarray(factors, FACTOR_DTYPE)
    qt = np.ascontiguousarray(qt, np.float64)
    cost = C.c_double()
    H = np.zeros((6, 6))
    g = np.zeros(6)
    res = np.zeros((len(f), 3)) if want_residuals else None
    lib().orc_eval(_p(f), len(f), _p(qt), C.c_double(huber_a), C.byref(cost), _p(H), _p(g),
                   _p(res) if want_residuals else None)
    return (cost.value, H, g, res) if want_residuals else (cost.value, H, g)


def solve(factors, qt, max_iter=4, huber_a=0.1):
    f = np.ascontiguousarray(factors, FACTOR_DTYPE)
    x = np.array(qt, np.float64)
    s = SolveSummary()
    lib().orc_solve(_p(f), len(f), _p(x), max_iter, C.c_double(huber_a), C.byref(s))
    return x, s


_lib_nf = None


def lib_nanoflann():
    """The same restatement compiled on the reference's vendored nanoflann tree (oracle/_ref/libref_aloam.so): the k-d
    tree of the CPU baseline ("Ceres + kd-tree").  None when it was never built."""
    global _lib_nf
    if _lib_nf is None:
        if not os.path.exists(_REF_ALOAM):
            build()
        if not os.path.exists(_REF_ALOAM):
            return None
        _lib_nf = C.CDLL(_REF_ALOAM)
    return _lib_nf


def register_aloam(map_corner, map_surf, corner, surf, qt, outer=2, max_iter=4, tree="own"):
    """tree = "own": the oracle's private k-d tree; "nanoflann": the reference's vendored nanoflann (CPU baseline)."""
    mc, ms, c, s = _f32(map_corner), _f32(map_surf), _f32(corner), _f32(surf)
    x = np.array(qt, np.float64)
    sums = (SolveSummary * outer)()
    nf = np.zeros(2 * outer, np.int32)
    L = lib() if tree == "own" else lib_nanoflann()
    if L is None:
        raise RuntimeError("oracle/_ref/libref_aloam.so not available")
    n = L.orc_register_aloam(_p(mc), len(mc), _p(ms), len(ms), _stride(mc), _p(c), len(c), _p(s), len(s),
                                 _stride(c) if len(c) else _stride(s), _p(x), outer, max_iter, sums, _p(nf))
    return x, list(sums)[:n], nf


# ------------------------------------------------------------------------------------------------
# front end
# ------------------------------------------------------------------------------------------------
class FeatureCounts(C.Structure):
    _fields_ = [("n_cloud", C.c_int32), ("n_sharp", C.c_int32), ("n_less_sharp", C.c_int32), ("n_flat", C.c_int32),
                ("n_less_flat", C.c_int32), ("ring_start", C.c_int32 * 64), ("ring_end", C.c_int32 * 64)]


def _ioff(a):
    return 4 if a.shape[1] >= 8 else 3


def project(cloud, H=64, W=1024):
    """ImageHandler::cloud_handler: (image_range u8, image_intensity u8, cloud_track (H*W,4) f32)."""
    c = _f32(cloud)
    assert len(c) == H * W
    rng = np.zeros((H, W), np.uint8)
    inten = np.zeros((H, W), np.uint8)
    track = np.zeros((H * W, 4), np.float32)
    lib().orc_project(_p(c), H, W, c.strides[0], _ioff(c), _p(rng), _p(inten), _p(track))
    return rng, inten, track


def voxelgrid(cloud, leaf):
    """pcl::VoxelGrid (PCL 1.10 restated): centroids (m,4) xyzi in ascending voxel index."""
    c = _f32(cloud)
    out = np.zeros((max(len(c), 1), 4), np.float32)
    if len(c) == 0:
        return out[:0]
    m = lib().orc_voxelgrid(_p(c), len(c), c.strides[0], _ioff(c), C.c_float(leaf), _p(out))
    return out[:m].copy()


def extract_features(cloud, min_range=0.3):
    """scanRegistration.cpp laserCloudHandler (64-ring path).  Returns a dict of arrays."""
    c = _f32(cloud)
    n = len(c)
    cl = np.zeros((n, 4), np.float32)
    curv = np.zeros(n, np.float32)
    label = np.zeros(n, np.int32)
    src = np.zeros(n, np.int32)
    sharp = np.zeros(n, np.int32)
    lsharp = np.zeros(n, np.int32)
    flat = np.zeros(n, np.int32)
    lflat = np.zeros((n, 4), np.float32)
    cnt = FeatureCounts()
    lib().orc_extract_features(_p(c), n, c.strides[0], C.c_float(min_range), _p(cl), _p(curv), _p(label), _p(src),
                               _p(sharp), _p(lsharp), _p(flat), _p(lflat), C.byref(cnt))
    N = cnt.n_cloud
    return dict(cloud=cl[:N], curvature=curv[:N], label=label[:N], src_index=src[:N], sharp_idx=sharp[:cnt.n_sharp],
                less_sharp_idx=lsharp[:cnt.n_less_sharp], flat_idx=flat[:cnt.n_flat], less_flat=lflat[:cnt.n_less_flat],
                ring_start=np.array(cnt.ring_start[:]), ring_end=np.array(cnt.ring_end[:]))


# ------------------------------------------------------------------------------------------------
# ScanContext
# ------------------------------------------------------------------------------------------------
def sc_make(points):
    """SCManager::makeScancontext -> (20, 60) float64 (values are floats widened to double)."""
    p = _f32(points)
    d = np.zeros((20, 60), np.float64)
    lib().orc_sc_make(_p(p), len(p), _stride(p), _p(d))
    return d


def sc_keys(desc):
    d = np.ascontiguousarray(desc, np.float64)
    rk, sk = np.zeros(20), np.zeros(60)
    lib().orc_sc_keys(_p(d), _p(rk), _p(sk))
    return rk, sk


def sc_distance(q, c):
    q = np.ascontiguousarray(q, np.float64)
    c = np.ascontiguousarray(c, np.float64)
    d = C.c_double()
    s = C.c_int32()
    lib().orc_sc_distance(_p(q), _p(c), C.byref(d), C.byref(s))
    return d.value, s.value


def sc_topk(db, q, k=10):
    db = np.ascontiguousarray(db, np.float64).reshape(-1, 1200)
    q = np.ascontiguousarray(q, np.float64)
    dist = np.zeros(k)
    idx = np.zeros(k, np.int32)
    sh = np.zeros(k, np.int32)
    lib().orc_sc_topk(_p(db), len(db), _p(q), k, _p(dist), _p(idx), _p(sh))
    return dist, idx, sh


def sc_detect_loop_reference(db, q, num_candidates=10):
    """detectLoopClosureID (Scancontext.cpp:253-344) at a tree-refresh boundary: 10-NN of the float ring key in the
    reference's own nanoflann tree (oracle/_ref), then the best candidate by distanceBtnScanContext.
    Returns (nn_idx, min_dist, nn_align, candidate ids)."""
    r = ref()
    if r is None:
        raise RuntimeError("oracle/_ref not built")
    db = np.ascontiguousarray(db, np.float64).reshape(-1, 20, 60)
    keys = np.stack([sc_keys(d)[0] for d in db]).astype(np.float32)  # eig2stdvec: double -> float
    qk = sc_keys(q)[0].astype(np.float32)
    idx = np.zeros(num_candidates, np.uint64)
    d2 = np.zeros(num_candidates, np.float32)
    r.ref_ringkey_knn.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_void_p, C.c_int, C.c_void_p, C.c_void_p]
    r.ref_ringkey_knn(_p(keys), len(keys), 20, _p(qk), num_candidates, _p(idx), _p(d2))
    best, arg, align = 10000000.0, 0, 0
    for ci in idx:
        d, s = sc_distance(q, db[int(ci)])
        if d < best:
            best, arg, align = d, int(ci), s
    return arg, best, align, idx.astype(np.int64)


# ------------------------------------------------------------------------------------------------
# scan-to-scan odometry
# ------------------------------------------------------------------------------------------------
def odom_associate(last_corner, last_surf, sharp, flat, qt):
    lc, ls, sh, fl = _f32(last_corner), _f32(last_surf), _f32(sharp), _f32(flat)
    assert lc.shape[1] >= 4 and _stride(lc) == _stride(ls)
    qt = np.ascontiguousarray(qt, np.float64)
    out = np.zeros(len(sh) + len(fl), FACTOR_DTYPE)
    n = lib().orc_odom_associate(_p(lc), len(lc), _p(ls), len(ls), _stride(lc), _ioff(lc), _p(sh), len(sh), _p(fl), len(fl),
                                 _stride(sh) if len(sh) else _stride(fl), _p(qt), _p(out))
    return out[:n]


def odometry(last_corner, last_surf, sharp, flat, qt, outer=2, max_iter=4):
    lc, ls, sh, fl = _f32(last_corner), _f32(last_surf), _f32(sharp), _f32(flat)
    x = np.array(qt, np.float64)
    sums = (SolveSummary * outer)()
    nf = np.zeros(2 * outer, np.int32)
    lib().orc_odometry(_p(lc), len(lc), _p(ls), len(ls), _stride(lc), _ioff(lc), _p(sh), len(sh), _p(fl), len(fl),
                       _stride(sh) if len(sh) else _stride(fl), _p(x), outer, max_iter, sums, _p(nf))
    return x, list(sums), nf


# ------------------------------------------------------------------------------------------------
# ikd-Tree Add_Points
# ------------------------------------------------------------------------------------------------
def ikd_add_points(existing_xyz, add_xyz, ds, downsample=True):
    e = np.ascontiguousarray(np.asarray(existing_xyz, np.float32)[:, :3]).reshape(-1, 3)
    a = np.ascontiguousarray(np.asarray(add_xyz, np.float32)[:, :3]).reshape(-1, 3)
    out = np.zeros((len(e) + len(a) + 1, 3), np.float32)
    n = lib().orc_ikd_add_points(_p(e), len(e), _p(a), len(a), C.c_float(ds), 1 if downsample else 0, _p(out), len(out))
    return out[:n].copy()


_ref_fun = None


def ref_functors():
    """The reference's own Ceres cost functors on dual numbers, oracle/_ref/libref_functors.so (None when never built)."""
    global _ref_fun
    if _ref_fun is None:
        if not os.path.exists(_REF_FUN):
            build()
        if not os.path.exists(_REF_FUN):
            return None
        r = C.CDLL(_REF_FUN)
        r.ref_functor_eval.argtypes = [C.c_int] + [C.c_void_p] * 4 + [C.c_double] + [C.c_void_p] * 4
        _ref_fun = r
    return _ref_fun


def ref_functor_eval(ftype, p, a, b=(0, 0, 0), c=(0, 0, 0), s=1.0, qt=(0, 0, 0, 1, 0, 0, 0)):
    """One reference functor at pose qt = (qx, qy, qz, qw, tx, ty, tz): (residuals (R,), ambient Jacobian (R, 7)).
    ftype: 1 LidarEdgeFactor(p, a, b, s); 2 LidarPlaneNormFactor(p, n = a, d = b[0]); 3 front_end_residual(src = p,
    dst = a); 4 LidarPlaneFactor(p, j = a, l = b, m = c, s)."""
    arr = [np.ascontiguousarray(v, np.float64) for v in (p, a, b, c)]
    qt = np.ascontiguousarray(qt, np.float64)
    q, t = qt[:4].copy(), qt[4:].copy()
    r = np.zeros(3)
    J = np.zeros((3, 7))
    n = ref_functors().ref_functor_eval(int(ftype), _p(arr[0]), _p(arr[1]), _p(arr[2]), _p(arr[3]), C.c_double(s), _p(q), _p(t),
                                        _p(r), _p(J))
    return r[:n].copy(), J[:n].copy()


_ref_scanreg = None


def ref_scanreg():
    """The reference's own LOAM front end (src/scanRegistration.cpp:227-589 cut out of its ROS node),
    oracle/_ref/libref_scanreg.so (None when never built)."""
    global _ref_scanreg
    if _ref_scanreg is None:
        if not os.path.exists(_REF_SCANREG):
            build()
        if not os.path.exists(_REF_SCANREG):
            return None
        r = C.CDLL(_REF_SCANREG)
        r.ref_scanreg.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_double, C.c_int] + [C.c_void_p] * 11
        _ref_scanreg = r
    return _ref_scanreg


def ref_extract_features(cloud, min_range=0.3, n_scans=64):
    """The reference code's laserCloudHandler numeric body on `cloud`: dict with the ring-ordered cloud, curvature, label,
    the four feature clouds (xyzi rows; less_flat after the oracle's VoxelGrid stand-in, less_flat_raw before it) and
    scanStartInd / scanEndInd."""
    c = _f32(cloud)
    n = len(c)
    cap = n + 8
    bufs = {k: np.zeros((cap, 4), np.float32) for k in ("cloud", "sharp", "less_sharp", "flat", "less_flat", "less_flat_raw")}
    curv = np.zeros(cap, np.float32)
    label = np.zeros(cap, np.int32)
    rs, re_ = np.zeros(64, np.int32), np.zeros(64, np.int32)
    cnt = np.zeros(8, np.int32)
    r = ref_scanreg()
    if r is None:
        raise RuntimeError("oracle/_ref/libref_scanreg.so not available")
    rc = r.ref_scanreg(_p(c), n, c.strides[0], int(n_scans), C.c_double(min_range), cap, _p(bufs["cloud"]), _p(curv), _p(label),
                       _p(bufs["sharp"]), _p(bufs["less_sharp"]), _p(bufs["flat"]), _p(bufs["less_flat"]), _p(bufs["less_flat_raw"]),
                       _p(rs), _p(re_), _p(cnt))
    if rc != 0:
        raise RuntimeError("ref_scanreg failed")
    out = {k: bufs[k][:cnt[i]].copy() for i, k in enumerate(("cloud", "sharp", "less_sharp", "flat", "less_flat", "less_flat_raw"))}
    out["curvature"], out["label"] = curv[:cnt[0]].copy(), label[:cnt[0]].copy()
    out["ring_start"], out["ring_end"] = rs, re_
    return out


_ref_laserodom = None


def ref_laserodom():
    """The reference's own scan-to-scan association (src/laserOdometry.cpp:417-713 cut out of its node),
    oracle/_ref/libref_laserodom.so (None when never built)."""
    global _ref_laserodom
    if _ref_laserodom is None:
        if not os.path.exists(_REF_LASERODOM):
            build()
        if not os.path.exists(_REF_LASERODOM):
            return None
        r = C.CDLL(_REF_LASERODOM)
        r.ref_laserodom_associate.argtypes = [C.c_void_p, C.c_int, C.c_void_p, C.c_int, C.c_void_p, C.c_int, C.c_void_p, C.c_int] + [C.c_void_p] * 5
        _ref_laserodom = r
    return _ref_laserodom


def ref_odom_associate(last_corner, last_surf, sharp, flat, qt):
    """The residual blocks the reference code builds at pose qt = (qx, qy, qz, qw, tx, ty, tz): (edge (Ne, 10) = curr 3,
    last_point_a 3, last_point_b 3, s;  plane (Np, 13) = curr 3, last_point_j / l / m 3 each, s;  the reference's own
    (corner_correspondence, plane_correspondence) counters)."""
    r = ref_laserodom()
    if r is None:
        raise RuntimeError("oracle/_ref/libref_laserodom.so not available")
    cl = [np.ascontiguousarray(np.asarray(a, np.float32)[:, :4]) for a in (last_corner, last_surf, sharp, flat)]
    qt = np.ascontiguousarray(qt, np.float64)
    q, t = qt[:4].copy(), qt[4:].copy()
    edge = np.zeros((len(cl[2]) + 1, 10))
    plane = np.zeros((len(cl[3]) + 1, 13))
    cnt = np.zeros(4, np.int32)
    rc = r.ref_laserodom_associate(_p(cl[0]), len(cl[0]), _p(cl[1]), len(cl[1]), _p(cl[2]), len(cl[2]), _p(cl[3]), len(cl[3]), _p(q), _p(t),
                                   _p(edge), _p(plane), _p(cnt))
    if rc != 0:
        raise RuntimeError(f"ref_laserodom_associate failed ({rc})")
    return edge[:cnt[0]].copy(), plane[:cnt[1]].copy(), (int(cnt[2]), int(cnt[3]))


_ref_lasermapping = None


def ref_lasermapping():
    """The reference's own map maintenance (src/laserMapping.cpp:327-623, 875-945, 984-1004 cut out of process()),
    oracle/_ref/libref_lasermapping.so (None when never built)."""
    global _ref_lasermapping
    if _ref_lasermapping is None:
        if not os.path.exists(_REF_LASERMAPPING):
            build()
        if not os.path.exists(_REF_LASERMAPPING):
            return None
        r = C.CDLL(_REF_LASERMAPPING)
        r.ref_lasermapping_reset.argtypes = [C.c_float, C.c_float]
        r.ref_lasermapping_frame.argtypes = [C.c_void_p, C.c_int, C.c_void_p, C.c_int] + [C.c_void_p] * 7
        r.ref_lasermapping_cube.argtypes = [C.c_int, C.c_int, C.c_void_p, C.c_int]
        r.ref_lasermapping_associate.argtypes = [C.c_void_p, C.c_int, C.c_void_p, C.c_int, C.c_void_p, C.c_int, C.c_void_p, C.c_int] + [C.c_void_p] * 4
        _ref_lasermapping = r
    return _ref_lasermapping


class RefLaserMapping:
    """The reference code's cube map (global state of the compiled fragment: one instance at a time).  frame() = one
    process() iteration WITHOUT the optimisation block: (pose used for the insertion (7,), window centre (3,), valid cube
    indices in the reference's order, (map corner, map surf, stack corner, stack surf) sizes)."""

    def __init__(self, line_res=0.4, plane_res=0.8):
        self._r = ref_lasermapping()
        if self._r is None:
            raise RuntimeError("oracle/_ref/libref_lasermapping.so not available")
        self._r.ref_lasermapping_reset(line_res, plane_res)

    def frame(self, corner_last, surf_last, qt_odom):
        c, s = CubeMap._x4(corner_last), CubeMap._x4(surf_last)
        qt = np.ascontiguousarray(qt_odom, np.float64)
        q, t = qt[:4].copy(), qt[4:].copy()
        out, cen, nv, valid, sizes = np.zeros(7), np.zeros(3, np.int32), np.zeros(1, np.int32), np.zeros(125, np.int32), np.zeros(4, np.int32)
        self._r.ref_lasermapping_frame(_p(c), len(c), _p(s), len(s), _p(q), _p(t), _p(out), _p(cen), _p(nv), _p(valid), _p(sizes))
        return out, cen, valid[:nv[0]].copy(), sizes

    def cube(self, which, index):
        n = self._r.ref_lasermapping_cube(which, index, None, 0)
        out = np.zeros((max(n, 1), 4), np.float32)
        self._r.ref_lasermapping_cube(which, index, _p(out), n)
        return out[:n]


def ref_map_associate(map_corner, map_surf, stack_corner, stack_surf, qt):
    """The residual blocks the reference's scan-to-map association (laserMapping.cpp:624-873) builds at pose qt: (edge (Ne, 10)
    = curr 3, point_a 3, point_b 3, s;  plane (Np, 7) = curr 3, unit normal 3, negative_OA_dot_norm)."""
    r = ref_lasermapping()
    if r is None:
        raise RuntimeError("oracle/_ref/libref_lasermapping.so not available")
    cl = [CubeMap._x4(a) for a in (map_corner, map_surf, stack_corner, stack_surf)]
    qt = np.ascontiguousarray(qt, np.float64)
    edge, plane, cnt = np.zeros((len(cl[2]) + 1, 10)), np.zeros((len(cl[3]) + 1, 7)), np.zeros(2, np.int32)
    rc = r.ref_lasermapping_associate(_p(cl[0]), len(cl[0]), _p(cl[1]), len(cl[1]), _p(cl[2]), len(cl[2]), _p(cl[3]), len(cl[3]), _p(qt),
                                      _p(edge), _p(plane), _p(cnt))
    if rc != 0:
        raise RuntimeError(f"ref_lasermapping_associate failed ({rc})")
    return edge[:cnt[0]].copy(), plane[:cnt[1]].copy()


_ref_scancontext = None


def ref_scancontext():
    """The reference's own ScanContext code (src/Scancontext.cpp + include/Scancontext.h, compiled unmodified),
    oracle/_ref/libref_scancontext.so (None when never built)."""
    global _ref_scancontext
    if _ref_scancontext is None:
        if not os.path.exists(_REF_SCANCONTEXT):
            build()
        if not os.path.exists(_REF_SCANCONTEXT):
            return None
        r = C.CDLL(_REF_SCANCONTEXT)
        r.ref_sc_make.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_void_p]
        r.ref_sc_keys.argtypes = [C.c_void_p] * 3
        r.ref_sc_distance.argtypes = [C.c_void_p] * 4
        r.ref_sc_detect.argtypes = [C.c_void_p, C.c_int, C.c_void_p]
        _ref_scancontext = r
    return _ref_scancontext


def ref_sc_make(points):
    p = _f32(points)
    d = np.zeros((20, 60), np.float64)
    ref_scancontext().ref_sc_make(_p(p), len(p), _stride(p), _p(d))
    return d


def ref_sc_keys(desc):
    d = np.ascontiguousarray(desc, np.float64)
    rk, sk = np.zeros(20), np.zeros(60)
    ref_scancontext().ref_sc_keys(_p(d), _p(rk), _p(sk))
    return rk, sk


def ref_sc_distance(q, c):
    a, b = np.ascontiguousarray(q, np.float64), np.ascontiguousarray(c, np.float64)
    dist, shift = np.zeros(1), np.zeros(1, np.int32)
    ref_scancontext().ref_sc_distance(_p(a), _p(b), _p(dist), _p(shift))
    return float(dist[0]), int(shift[0])


def ref_sc_detect(db, q):
    """SCManager::detectLoopClosureID with the keyframes db[0..n) saved first and q saved last (the query), through the
    reference's own code end to end (nanoflann tree included): (loop id or -1, yaw difference in rad)."""
    d = np.ascontiguousarray(np.concatenate([np.asarray(db, np.float64).reshape(-1, 1200), np.asarray(q, np.float64).reshape(1, 1200)]))
    yaw = np.zeros(1, np.float32)
    lid = ref_scancontext().ref_sc_detect(_p(d), len(d), _p(yaw))
    return int(lid), float(yaw[0])


_ref_imagehandler = None


def ref_imagehandler():
    """The reference's own projection loop (src/image_handler.h_ouster:113-139), oracle/_ref/libref_imagehandler.so
    (None when never built)."""
    global _ref_imagehandler
    if _ref_imagehandler is None:
        if not os.path.exists(_REF_IMAGEHANDLER):
            build()
        if not os.path.exists(_REF_IMAGEHANDLER):
            return None
        r = C.CDLL(_REF_IMAGEHANDLER)
        r.ref_project.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p]
        _ref_imagehandler = r
    return _ref_imagehandler


def ref_project(cloud, H=64, W=1024):
    """ImageHandler::cloud_handler's loop through the reference code: (image_range u8, image_intensity u8, cloud_track)."""
    c = np.ascontiguousarray(np.asarray(cloud, np.float32)[:, :4])
    assert len(c) == H * W
    rng, inten, track = np.zeros((H, W), np.uint8), np.zeros((H, W), np.uint8), np.zeros((H * W, 4), np.float32)
    ref_imagehandler().ref_project(_p(c), H, W, c.strides[0], _p(rng), _p(inten), _p(track))
    return rng, inten, track


_ref_ikd = None


def ref_ikd():
    """The reference's own ikd-Tree, oracle/_ref/libref_ikd.so (None when it was never built)."""
    global _ref_ikd
    if _ref_ikd is None:
        if not os.path.exists(_REF_IKD):
            build()
        if not os.path.exists(_REF_IKD):
            return None
        r = C.CDLL(_REF_IKD)
        r.ref_ikd_create.restype = C.c_void_p
        r.ref_ikd_create.argtypes = [C.c_float, C.c_float, C.c_float]
        r.ref_ikd_free.argtypes = [C.c_void_p]
        r.ref_ikd_size.argtypes = [C.c_void_p]
        r.ref_ikd_build.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_int]
        r.ref_ikd_nearest.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p]
        r.ref_ikd_add_points.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int]
        r.ref_ikd_flatten.argtypes = [C.c_void_p, C.c_void_p, C.c_int]
        _ref_ikd = r
    return _ref_ikd


class RefIkdTree:
    """KD_TREE<pcl::PointXYZ> of the reference (src/ikd-Tree), the way mapOptimization.cpp:504 constructs it:
    (delete criterion 0.3, balance criterion 0.6, down-sample box 0.4)."""

    def __init__(self, delete_param=0.3, balance_param=0.6, box_length=0.4):
        self._r = ref_ikd()
        if self._r is None:
            raise RuntimeError("oracle/_ref/libref_ikd.so not available")
        self._h = self._r.ref_ikd_create(delete_param, balance_param, box_length)

    def build(self, xyz):  # Build(points)  mapOptimization.cpp:192
        a = _f32(xyz)
        self._r.ref_ikd_build(self._h, _p(a), len(a), _stride(a))
        return self

    def size(self):
        return self._r.ref_ikd_size(self._h)

    def nearest(self, q_xyz, k=5):  # Nearest_Search  mapOptimization.cpp:393 -> (points nq x k x 3, d2 nq x k, found nq)
        q = _f32(q_xyz)
        pts = np.full((len(q), k, 3), np.nan, np.float32)
        d2 = np.full((len(q), k), np.inf, np.float32)
        cnt = np.zeros(len(q), np.int32)
        self._r.ref_ikd_nearest(self._h, _p(q), len(q), _stride(q), k, _p(pts), _p(d2), _p(cnt))
        return pts, d2, cnt

    def add_points(self, xyz, downsample=True):  # Add_Points(points, true)  mapOptimization.cpp:475
        a = _f32(xyz)
        return self._r.ref_ikd_add_points(self._h, _p(a), len(a), _stride(a), 1 if downsample else 0)

    def points(self):  # flatten(Root_Node, storage, NOT_RECORD): tree order
        n = self.size()
        out = np.zeros((n + 16, 3), np.float32)
        m = self._r.ref_ikd_flatten(self._h, _p(out), len(out))
        return out[:m].copy()

    def close(self):
        if self._h:
            self._r.ref_ikd_free(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


# ------------------------------------------------------------------------------------------------
# laserMapping rolling cube map
# ------------------------------------------------------------------------------------------------
class CubeFrameStats(C.Structure):
    _fields_ = [("n_map_corner", C.c_int32), ("n_map_surf", C.c_int32), ("n_stack_corner", C.c_int32),
                ("n_stack_surf", C.c_int32), ("ran_optimization", C.c_int32), ("n_valid", C.c_int32),
                ("cen", C.c_int32 * 3), ("pad", C.c_int32)]


class CubeMap:
    """laserMapping.cpp process() hot section (rolling 21x21x11 cube map + guarded registration + insertion)."""

    def __init__(self, line_res=0.4, plane_res=0.8):
        l = lib()
        l.orc_cubemap_create.restype = C.c_void_p
        l.orc_cubemap_create.argtypes = [C.c_float, C.c_float]
        l.orc_cubemap_destroy.argtypes = [C.c_void_p]
        l.orc_cubemap_insert_world.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_void_p, C.c_int, C.c_void_p]
        l.orc_cubemap_frame.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_void_p, C.c_int, C.c_void_p, C.c_void_p,
                                        C.c_void_p, C.c_void_p]
        l.orc_cubemap_cube.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_void_p, C.c_int]
        self._l = l
        self._h = l.orc_cubemap_create(line_res, plane_res)

    def __del__(self):
        try:
            self._l.orc_cubemap_destroy(self._h)
        except Exception:
            pass

    @staticmethod
    def _x4(a):
        a = np.asarray(a, np.float32)
        out = np.zeros((len(a), 4), np.float32)
        out[:, :min(4, a.shape[1])] = a[:, :4]
        return out

    def set_solve(self, enabled: bool):
        """False: frames keep the pose transformAssociateToMap predicts (map-logic comparisons against the reference code)."""
        self._l.orc_cubemap_set_solve.argtypes = [C.c_void_p, C.c_int]
        self._l.orc_cubemap_set_solve(self._h, 1 if enabled else 0)

    def insert_world(self, corner, surf, centre):
        c, s = self._x4(corner), self._x4(surf)
        ctr = np.ascontiguousarray(centre, np.float64)
        self._l.orc_cubemap_insert_world(self._h, _p(c), len(c), _p(s), len(s), _p(ctr))

    def frame(self, corner_last, surf_last, qt_odom):
        c, s = self._x4(corner_last), self._x4(surf_last)
        qo = np.ascontiguousarray(qt_odom, np.float64)
        out = np.zeros(7)
        sums = (SolveSummary * 2)()
        st = CubeFrameStats()
        self._l.orc_cubemap_frame(self._h, _p(c), len(c), _p(s), len(s), _p(qo), _p(out), sums, C.byref(st))
        return out, list(sums), st

    def cube(self, which, index, cap=1 << 16):
        out = np.zeros((cap, 4), np.float32)
        n = self._l.orc_cubemap_cube(self._h, which, index, _p(out), cap)
        return out[:n].copy()


# ------------------------------------------------------------------------------------------------
# the three nodes chained (scanRegistration -> laserOdometry -> laserMapping), one frame per call
# ------------------------------------------------------------------------------------------------
def _qmul(a, b):  # Eigen quaternion product, (x, y, z, w)
    return np.array([a[3] * b[0] + a[0] * b[3] + a[1] * b[2] - a[2] * b[1],
                     a[3] * b[1] + a[1] * b[3] + a[2] * b[0] - a[0] * b[2],
                     a[3] * b[2] + a[2] * b[3] + a[0] * b[1] - a[1] * b[0],
                     a[3] * b[3] - a[0] * b[0] - a[1] * b[1] - a[2] * b[2]])


def _qrot(q, v):  # Eigen QuaternionBase::_transformVector
    u = q[:3]
    uv = 2.0 * np.cross(u, v)
    return v + q[3] * uv + np.cross(u, uv)


class Slam:
    """laserOdometry.cpp main loop (:376-845) between the front end and CubeMap.frame, restated with the oracle's own
    pieces: first frame only initialises; afterwards 2 x (association + Solve) from the previous q/t_last_curr (when
    use_aloam), t_w_curr += q_w_curr * t_last_curr, q_w_curr *= q_last_curr (:716-717), last clouds <- less-sharp /
    less-flat (:793-808), then one process() iteration of laserMapping with the odometry pose."""

    def __init__(self, line_res=0.4, plane_res=0.8, min_range=0.3, mapping="laserMapping"):
        self.cube = CubeMap(line_res, plane_res) if mapping == "laserMapping" else None
        self.mapopt = None if mapping == "laserMapping" else MapOptimization()
        self.min_range = min_range
        self.inited = False
        self.para = np.array([0, 0, 0, 1, 0, 0, 0.0])
        self.q_w_curr = np.array([0, 0, 0, 1.0])
        self.t_w_curr = np.zeros(3)
        self.last_corner = self.last_surf = None

    def frame(self, cloud, use_aloam=True):
        f = extract_features(cloud, self.min_range)
        sharp, flat = f["cloud"][f["sharp_idx"]], f["cloud"][f["flat_idx"]]
        lsharp, lflat = f["cloud"][f["less_sharp_idx"]], f["less_flat"]
        odo = None
        if not self.inited:
            self.inited = True
        else:
            if use_aloam:
                self.para, sums, nf = odometry(self.last_corner, self.last_surf, sharp, flat, self.para)
                odo = (sums, nf)
            self.t_w_curr = self.t_w_curr + _qrot(self.q_w_curr, self.para[4:])
            self.q_w_curr = _qmul(self.q_w_curr, self.para[:4])
        self.last_corner, self.last_surf = lsharp.copy(), lflat.copy()
        qt_odom = np.concatenate([self.q_w_curr, self.t_w_curr])
        if self.cube is not None:
            qt_map, msums, st = self.cube.frame(lsharp, lflat, qt_odom)
            return qt_odom, qt_map, dict(features=f, odometry=odo, mapping=msums, cubemap=st)
        qt_map, info = self.mapopt.frame(cloud, lflat, self.q_w_curr, self.t_w_curr)
        return qt_odom, qt_map, dict(features=f, odometry=odo, mapopt=info)


# ------------------------------------------------------------------------------------------------
# intensity-image feature matching back end (intensity_feature_tracker.cpp:631-738, 880-928)
# ------------------------------------------------------------------------------------------------
_POPC = np.array([bin(i).count("1") for i in range(256)], np.uint16)


def hamming_matrix(a, b):
    """All-pairs Hamming distances of two uint8 descriptor sets (n1 x 32, n2 x 32)."""
    a = np.ascontiguousarray(a, np.uint8)
    b = np.ascontiguousarray(b, np.uint8)
    out = np.zeros((len(a), len(b)), np.int32)
    for i in range(len(a)):  # row by row keeps memory small
        out[i] = _POPC[np.bitwise_xor(b, a[i])].sum(axis=1)
    return out


def bf_match_hamming(cur, prev, cross_check=True):
    """cv::BFMatcher(NORM_HAMMING, crossCheck).match(cur, prev) restated (OpenCV batchDistance, K = 1): every query
    takes its nearest train descriptor, first minimum wins (lowest train index); with crossCheck the pair is kept only
    if the query is also the nearest (first minimum = lowest query index) of that train descriptor.
    Returns (queryIdx, trainIdx, distance) in query order."""
    if len(cur) == 0 or len(prev) == 0:
        z = np.zeros(0, np.int32)
        return z, z.copy(), np.zeros(0, np.float32)
    d = hamming_matrix(cur, prev)
    best_t = d.argmin(axis=1)  # numpy argmin returns the first minimum
    keep = np.ones(len(cur), bool)
    if cross_check:
        best_q = d.argmin(axis=0)
        keep = best_q[best_t] == np.arange(len(cur))
    qi = np.nonzero(keep)[0].astype(np.int32)
    return qi, best_t[qi].astype(np.int32), d[qi, best_t[qi]].astype(np.float32)


def good_matches(qi, ti, dist, fraction=0.3):
    """std::sort(matches) + the first `i < matches.size() * fraction` (intensity_feature_tracker.cpp:643-648); DMatch
    orders by distance only, equal distances keep query order here (the declared tie-break)."""
    order = np.lexsort((qi, dist))
    n_good = int(np.ceil(len(qi) * float(fraction)))
    sel = order[:n_good]
    return qi[sel], ti[sel], dist[sel]


def align_points(src_xyz, dst_xyz, qt0=(0, 0, 0, 1, 0, 0, 0), max_iter=20, huber_a=0.1):
    """p2p_calculateRandT (intensity_feature_tracker.cpp:880-928): front_end_residual blocks, HuberLoss(0.1),
    EigenQuaternionParameterization, LM <= 20 iterations from the identity."""
    src = np.asarray(src_xyz, np.float32)[:, :3]
    dst = np.asarray(dst_xyz, np.float32)[:, :3]
    f = np.zeros(len(src), FACTOR_DTYPE)
    f["type"] = 3
    f["src"] = np.arange(len(src))
    f["p"] = src.astype(np.float64)
    f["a"] = dst.astype(np.float64)
    return solve(f, np.array(qt0, np.float64), max_iter, huber_a)


# ------------------------------------------------------------------------------------------------
# ground-plane extraction (ImageHandler::groundPlaneExtraction, image_handler.h_ouster:41-100)
# ------------------------------------------------------------------------------------------------
def ground_sample_triples(m, count=64, seed=1):
    """The DECLARED sampler (PCL's rand()-based one is unpinned): LCG x <- 1664525 x + 1013904223 (mod 2^32),
    index = (x >> 8) mod m, three distinct indices per hypothesis."""
    x = np.uint64(seed & 0xFFFFFFFF)
    out = np.zeros((count, 3), np.int32)

    def nxt():
        nonlocal x
        x = (x * np.uint64(1664525) + np.uint64(1013904223)) & np.uint64(0xFFFFFFFF)
        return int(x >> np.uint64(8)) % m
    for h in range(count):
        a = nxt()
        b = nxt()
        while b == a:
            b = nxt()
        c = nxt()
        while c == a or c == b:
            c = nxt()
        out[h] = (a, b, c)
    return out


def _plane_from_triple(p0, p1, p2):
    """SampleConsensusModelPlane::computeModelCoefficients in float32, left-to-right, no FMA.  Returns (coeff[4], ok)."""
    f = np.float32
    a = (p1 - p0).astype(f)
    b = (p2 - p0).astype(f)
    with np.errstate(divide="ignore", invalid="ignore"):
        r = (a / b).astype(f)
    if r[0] == r[1] and r[2] == r[1]:
        return np.zeros(4, f), False
    n = np.array([f(a[1] * b[2]) - f(a[2] * b[1]), f(a[2] * b[0]) - f(a[0] * b[2]), f(a[0] * b[1]) - f(a[1] * b[0])], f)
    nn = np.sqrt(f(f(f(n[0] * n[0]) + f(n[1] * n[1])) + f(n[2] * n[2])))
    with np.errstate(divide="ignore", invalid="ignore"):
        n = (n / nn).astype(f)
    d = f(-1.0) * f(f(f(n[0] * p0[0]) + f(n[1] * p0[1])) + f(n[2] * p0[2]))
    return np.array([n[0], n[1], n[2], d], f), bool(np.all(np.isfinite(n)))


def ground_plane(cloud, z_min=-2.0, z_max=-0.45, dist_thresh=0.01, max_iterations=50, probability=0.99, seed=1,
                 band=0.03, max_angle_deg=15.0):
    """groundPlaneExtraction restated: z-band screening, RANSAC plane (PCL 1.10 RandomSampleConsensus::computeModel loop
    with the declared sampler; inlier test |a x + b y + c z + d| < threshold in float), least-squares refit of the best
    model's inliers (PCA, smallest eigenvector, oriented upward; sums in double -- PCL's float accumulation order is
    not reproducible in parallel), acceptance n . z > cos(15 deg), ground = points within `band` of the plane with z < 0
    (input order).  Returns (ground_xyz, coeff float32[4], info dict)."""
    f = np.float32
    pts = np.ascontiguousarray(np.asarray(cloud, np.float32)[:, :3])
    z = pts[:, 2].astype(np.float64)
    scr = pts[(z >= z_min) & (z <= z_max)]
    m = len(scr)
    info = dict(n_band=m, best=-1, n_best=0, iterations=0, accepted=False)
    empty = np.zeros((0, 3), np.float32)
    if m < 3:
        return empty, np.zeros(4, f), info
    n_hyp = 64
    tri = ground_sample_triples(m, n_hyp, seed)
    models, valid, counts = [], [], []
    thr = f(dist_thresh)
    for h in range(n_hyp):
        co, ok = _plane_from_triple(scr[tri[h, 0]], scr[tri[h, 1]], scr[tri[h, 2]])
        models.append(co)
        valid.append(ok)
        if ok:
            dd = ((f(co[0]) * scr[:, 0] + f(co[1]) * scr[:, 1]).astype(f) + f(co[2]) * scr[:, 2]).astype(f) + f(co[3])
            counts.append(int((np.abs(dd.astype(f)) < thr).sum()))
        else:
            counts.append(0)
    # replay of RandomSampleConsensus::computeModel (PCL 1.10 ransac.hpp:48-140)
    k, it, skipped, h = 1.0, 0, 0, 0
    n_best, best = 0, -1
    log_p = np.log(1.0 - probability)
    eps = np.finfo(np.float64).eps
    while it < k and skipped < max_iterations * 10 and h < n_hyp:
        if not valid[h]:
            skipped += 1
            h += 1
            continue
        if counts[h] > n_best:
            n_best, best = counts[h], h
            w = n_best / float(m)
            p_no = 1.0 - w ** 3
            p_no = min(max(p_no, eps), 1.0 - eps)
            k = log_p / np.log(p_no)
        it += 1
        h += 1
        if it > max_iterations:
            break
    info.update(best=best, n_best=n_best, iterations=it, counts=np.array(counts), triples=tri)
    if best < 0:
        return empty, np.zeros(4, f), info
    co = models[best]
    dd = ((f(co[0]) * scr[:, 0] + f(co[1]) * scr[:, 1]).astype(f) + f(co[2]) * scr[:, 2]).astype(f) + f(co[3])
    inl = scr[np.abs(dd.astype(f)) < thr].astype(np.float64)
    if len(inl) > 3:  # optimizeModelCoefficients
        c = inl.mean(axis=0)
        q = inl - c
        cov = q.T @ q / len(inl)
        wv, vv = np.linalg.eigh(cov)
        nrm = vv[:, 0]
        if nrm[2] < 0:
            nrm = -nrm
        co = np.array([nrm[0], nrm[1], nrm[2], -float(nrm @ c)], np.float64).astype(f)
    A, B, C_, D = [float(v) for v in co]
    info["coeff"] = co
    if not (f(C_) > np.cos(max_angle_deg * np.pi / 180.0)):
        return empty, co, info
    info["accepted"] = True
    P = pts.astype(np.float64)
    with np.errstate(invalid="ignore"):
        height = np.abs(A * P[:, 0] + B * P[:, 1] + C_ * P[:, 2] + D) / np.sqrt(A * A + B * B + C_ * C_)
        keep = (height <= band) & (P[:, 2] < 0.0)
    return pts[keep], co, info


# ------------------------------------------------------------------------------------------------
# mapOptimization::mapOptimizationCallback, LiDAR part (mapOptimization.cpp:99-500)
# ------------------------------------------------------------------------------------------------
def _quat_to_mat(q):  # Eigen::Quaterniond::toRotationMatrix
    x, y, z, w = q
    tx, ty, tz = 2 * x, 2 * y, 2 * z
    twx, twy, twz = tx * w, ty * w, tz * w
    txx, txy, txz, tyy, tyz, tzz = tx * x, ty * x, tz * x, ty * y, tz * y, tz * z
    return np.array([[1 - (tyy + tzz), txy - twz, txz + twy], [txy + twz, 1 - (txx + tzz), tyz - twx],
                     [txz - twy, tyz + twx, 1 - (txx + tyy)]])


def _transform_cloud(q, t, pts):
    """pcl::transformPointCloud with an Eigen::Matrix4d: double math, rows summed left to right, float store."""
    R = _quat_to_mat(q)
    p = np.asarray(pts, np.float32)[:, :3].astype(np.float64)
    out = np.empty((len(p), 3), np.float32)
    for a in range(3):
        out[:, a] = (((R[a, 0] * p[:, 0] + R[a, 1] * p[:, 1]) + R[a, 2] * p[:, 2]) + t[a]).astype(np.float32)
    return out


class MapOptimization:
    def __init__(self, voxel_leaf=0.8, downsample_size=0.4):
        self.leaf, self.ds = voxel_leaf, downsample_size
        self.map = None
        self.q_wmap_wodom = np.array([0, 0, 0, 1.0])
        self.t_wmap_wodom = np.zeros(3)

    def frame(self, cloud, plane_cloud, q_wodom, t_wodom, **ground_kw):
        g, coeff, ginfo = ground_plane(cloud, **ground_kw)
        merged = np.concatenate([g, np.asarray(plane_cloud, np.float32)[:, :3]]).astype(np.float32)
        q_wodom, t_wodom = np.asarray(q_wodom, np.float64), np.asarray(t_wodom, np.float64)
        qw = _qmul(self.q_wmap_wodom, q_wodom)
        tw = _qrot(self.q_wmap_wodom, t_wodom) + self.t_wmap_wodom
        info = dict(ground=ginfo, n_ground=len(g), ran_optimization=0, converged=False)
        if self.map is None:
            self.map = _transform_cloud(qw, tw, merged)
            info["map_size"] = len(self.map)
            return np.concatenate([qw, tw]), info
        m4 = np.zeros((len(merged), 4), np.float32)
        m4[:, :3] = merged
        stack = voxelgrid(m4, self.leaf)
        fac = associate(np.zeros((0, 3), np.float32), self.map, np.zeros((0, 4), np.float32), stack, np.concatenate([qw, tw]))
        x, sm = solve(fac, np.concatenate([qw, tw]), 10, 0.1)
        conv = sm.termination == 0
        qk, tk = (x[:4], x[4:]) if conv else (qw, tw)
        if conv:  # transformUpdate
            n2 = float(q_wodom @ q_wodom)
            qinv = np.array([-q_wodom[0], -q_wodom[1], -q_wodom[2], q_wodom[3]]) / n2
            self.q_wmap_wodom = _qmul(qk, qinv)
            self.t_wmap_wodom = tk - _qrot(self.q_wmap_wodom, t_wodom)
        world = _transform_cloud(qk, tk, stack)
        self.map = ikd_add_points(self.map, world, self.ds, True)
        info.update(ran_optimization=1, converged=bool(conv), n_query=len(stack), summary=sm, map_size=len(self.map),
                    n_plane_factors=int((fac["type"] == 2).sum()), key=np.concatenate([qk, tk]))
        return x, info


# ------------------------------------------------------------------------------------------------
# sensor_msgs/PointCloud2 blob -> packed xyzi: the field map pcl::fromROSMsg builds for PointXYZI
# (scanRegistration.cpp:235, image_handler.h_ouster:44,106): x, y, z FLOAT32 at their message offsets, intensity at its
# offset (FLOAT32 in the reference's Ouster driver; other numeric types converted to float as the library's extension).
# ------------------------------------------------------------------------------------------------
_PC2_DTYPES = {2: "<u1", 4: "<u2", 6: "<u4", 7: "<f4", 8: "<f8"}


def pc2_unpack(blob, point_step, off_x, off_y, off_z, off_intensity=-1, intensity_datatype=7):
    b = np.ascontiguousarray(blob, np.uint8).reshape(-1, point_step)
    n = len(b)

    def field(off, dt):
        w = np.dtype(dt).itemsize
        return np.ascontiguousarray(b[:, off:off + w]).view(dt).reshape(n)

    out = np.zeros((n, 4), np.float32)
    out[:, 0], out[:, 1], out[:, 2] = field(off_x, "<f4"), field(off_y, "<f4"), field(off_z, "<f4")
    if off_intensity >= 0:
        out[:, 3] = field(off_intensity, _PC2_DTYPES[intensity_datatype]).astype(np.float32)
    return out


def pc2_pack(xyzi, point_step, off_x, off_y, off_z, off_intensity=-1, intensity_datatype=7, fill=0xAB):
    """Test helper: the inverse (a message blob with the given layout, padding bytes = fill)."""
    a = np.asarray(xyzi, np.float32)
    n = len(a)
    b = np.full((n, point_step), fill, np.uint8)
    for k, off in enumerate((off_x, off_y, off_z)):
        b[:, off:off + 4] = np.ascontiguousarray(a[:, k]).view(np.uint8).reshape(n, 4)
    if off_intensity >= 0:
        dt = np.dtype(_PC2_DTYPES[intensity_datatype])
        b[:, off_intensity:off_intensity + dt.itemsize] = np.ascontiguousarray(a[:, 3].astype(dt)).view(np.uint8).reshape(n, dt.itemsize)
    return b.reshape(-1)

