"""Second, independent restatements (numpy / plain Python, written from the reference text, sharing no code with
oracle/*.cpp) of the parts of the oracle that no reference binary or library pins: projection, curvature, PCL VoxelGrid,
ScanContext distance, the ikd-Tree down-sampled insertion.  The C++ oracle must agree with them bit for bit (float paths)
or to 1e-12 (double paths whose summation order Eigen leaves open)."""
import math

import numpy as np
import pytest


# ---------------------------------------------------------------------------------------------- projection
def _project_numpy(cloud, H, W):
    """image_handler.h_ouster:103-140: range = sqrt(x^2+y^2+z^2) (float), u8(min(range*20, 255)), u8(min(I, 255)), cloud_track
    zeroed where range < 0.1."""
    x, y, z, it = (cloud[:, k].astype(np.float32) for k in range(4))
    rng = np.sqrt((x * x + y * y) + z * z, dtype=np.float32)
    r8 = np.minimum(rng * np.float32(20.0), np.float32(255.0)).astype(np.uint8)       # C++ float -> uchar truncation
    i8 = np.minimum(it, np.float32(255.0)).astype(np.uint8)
    track = cloud[:, :4].astype(np.float32).copy()
    track[:, 3] = np.minimum(it, np.float32(255.0))   # the CLAMPED intensity is what cloud_track keeps (:121,130)
    track[rng < np.float32(0.1)] = 0
    return r8.reshape(H, W), i8.reshape(H, W), track


def test_projection_vs_numpy(oracle_mod):
    rng = np.random.default_rng(0)
    H, W = 8, 64
    c = rng.normal(0, 6, (H * W, 4)).astype(np.float32)
    c[:, 3] = rng.uniform(0, 400, H * W)          # intensities above 255 saturate
    c[::7, :3] = 0                                # no-return rays
    c[1, :3] = [0.05, 0.02, 0.01]                 # below the 0.1 m cut
    c[2, :3] = [30, 40, 5]                        # range * 20 > 255 saturates
    r8, i8, tr = oracle_mod.project(c, H, W)
    w8, wi8, wtr = _project_numpy(c, H, W)
    assert np.array_equal(r8, w8) and np.array_equal(i8, wi8) and np.array_equal(tr, wtr)


# ---------------------------------------------------------------------------------------------- curvature
def test_curvature_vs_numpy(oracle_mod, ilsm):
    """scanRegistration.cpp:397-412: diff = sum of the 10 neighbours - 10 * p, added left to right in float; c = dx^2+dy^2+dz^2."""
    c = ilsm.synth.config1(n_map=20_000)
    fe = oracle_mod.extract_features(c["cloud"])
    p = fe["cloud"][:, :3].astype(np.float32)
    n = len(p)
    want = np.zeros(n, np.float32)
    order = [-5, -4, -3, -2, -1, 1, 2, 3, 4, 5]   # the reference's order of terms ...
    i = np.arange(5, n - 5)
    d = np.zeros((len(i), 3), np.float32)
    for o in order[:5]:
        d = d + p[i + o]
    d = d - np.float32(10) * p[i]                 # ... with "- 10 * p" in sixth place
    for o in order[5:]:
        d = d + p[i + o]
    want[5:n - 5] = (d[:, 0] * d[:, 0] + d[:, 1] * d[:, 1]) + d[:, 2] * d[:, 2]
    got = fe["curvature"]
    assert np.array_equal(got[5:n - 5], want[5:n - 5])


# ---------------------------------------------------------------------------------------------- VoxelGrid
def _voxelgrid_python(cloud, leaf):
    """PCL 1.10 VoxelGrid::applyFilter for a cubic leaf: min/max of the finite points, ijk = floor(p * inv) - min_b, index
    = i + j * dx + k * dx * dy, points grouped by index (stable: input order inside a voxel), float centroid of all four fields,
    output in ascending index."""
    f32 = np.float32
    inv = f32(1.0) / f32(leaf)
    pts = [r for r in cloud.astype(np.float32) if np.isfinite(r[:3]).all()]
    a = np.array(pts, np.float32)
    min_b = [int(math.floor(float(a[:, k].min() * inv))) for k in range(3)]
    max_b = [int(math.floor(float(a[:, k].max() * inv))) for k in range(3)]
    dx, dy = max_b[0] - min_b[0] + 1, max_b[1] - min_b[1] + 1
    groups = {}
    for r in pts:
        ijk = [int(f32(math.floor(float(r[k] * inv))) - f32(min_b[k])) for k in range(3)]
        groups.setdefault(ijk[0] + ijk[1] * dx + ijk[2] * dx * dy, []).append(r)
    out = []
    for key in sorted(groups):
        s = np.zeros(4, np.float32)
        for r in groups[key]:
            s = s + r[:4]
        out.append(s / f32(len(groups[key])))
    return np.array(out, np.float32)


@pytest.mark.parametrize("leaf", [0.2, 0.4, 0.8])
def test_voxelgrid_vs_python(oracle_mod, leaf):
    rng = np.random.default_rng(int(leaf * 10))
    c = rng.normal(0, 3, (1500, 4)).astype(np.float32)
    c[:, 2] *= 0.1
    c[:, 3] = rng.integers(0, 64, 1500) + 0.05
    c[100:140, :3] = c[60:100, :3] + rng.normal(0, 0.01, (40, 3)).astype(np.float32)  # several points per voxel
    got = oracle_mod.voxelgrid(c, leaf)
    want = _voxelgrid_python(c, leaf)
    assert got.shape == want.shape and len(want) < 1500
    assert np.array_equal(got, want)


# ---------------------------------------------------------------------------------------------- ScanContext distance
def _circshift(m, s):  # Scancontext.cpp:44-68: column c moves to (c + s) % cols
    return np.roll(m, s, axis=1)


def _dist_direct(a, b):  # :79-101
    num, tot = 0, 0.0
    for c in range(a.shape[1]):
        na, nb = np.linalg.norm(a[:, c]), np.linalg.norm(b[:, c])
        if na == 0 or nb == 0:
            continue
        tot += float(a[:, c] @ b[:, c]) / (na * nb)
        num += 1
    return 1.0 - tot / num


def _distance_btn(q, cand):  # :104-157
    k1, k2 = q.mean(axis=0)[None, :], cand.mean(axis=0)[None, :]
    best, arg = 10000000.0, 0
    for s in range(60):
        d = np.linalg.norm(k1 - _circshift(k2, s))
        if d < best:
            best, arg = d, s
    radius = int(round(0.5 * 0.1 * 60))
    space = sorted([arg] + [(arg + i + 60) % 60 for i in range(1, radius + 1)] + [(arg - i + 60) % 60 for i in range(1, radius + 1)])
    bd, bs = 10000000.0, 0
    for s in space:
        d = _dist_direct(q, _circshift(cand, s))
        if d < bd:
            bd, bs = d, s
    return bd, bs


def test_scancontext_distance_vs_numpy(oracle_mod, ilsm):
    db = ilsm.synth.sc_database(40).astype(np.float64)
    qs, ids, shifts = ilsm.synth.sc_queries(db.astype(np.float32), 6)
    for j in range(len(qs)):
        q = qs[j].astype(np.float64)
        for i in (int(ids[j]), (int(ids[j]) + 7) % 40):
            d, s = oracle_mod.sc_distance(q, db[i])
            wd, ws = _distance_btn(q, db[i])
            assert s == ws and abs(d - wd) < 1e-12
    # sector key / ring key
    rk, sk = oracle_mod.sc_keys(db[3])
    assert np.allclose(rk, db[3].mean(axis=1), rtol=0, atol=1e-15) and np.allclose(sk, db[3].mean(axis=0), rtol=0, atol=1e-15)


def test_scancontext_descriptor_vs_numpy(oracle_mod):
    """makeScancontext (:160-204): ring = ceil(r / 80 * 20), sector = ceil(theta / 360 * 60), max (z + 2) per bin."""
    rng = np.random.default_rng(4)
    p = np.zeros((4000, 3), np.float32)
    ang, rad = rng.uniform(0, 2 * np.pi, 4000), rng.uniform(0.5, 95, 4000)   # some beyond PC_MAX_RADIUS
    p[:, 0], p[:, 1], p[:, 2] = rad * np.cos(ang), rad * np.sin(ang), rng.uniform(-1.8, 5, 4000)
    want = np.full((20, 60), -1000.0)
    for x, y, z in p:
        zz = np.float32(np.float64(z) + 2.0)
        r = np.sqrt(np.float32(x * x + y * y), dtype=np.float32)
        if x >= 0 and y >= 0:
            th = (180 / np.pi) * np.float32(math.atan(np.float32(y / x)))
        elif x < 0 and y >= 0:
            th = 180 - (180 / np.pi) * np.float32(math.atan(np.float32(y / -x)))
        elif x < 0 and y < 0:
            th = 180 + (180 / np.pi) * np.float32(math.atan(np.float32(y / x)))
        else:
            th = 360 - (180 / np.pi) * np.float32(math.atan(np.float32(-y / x)))
        th = np.float32(th)
        if float(r) > 80.0:
            continue
        ring = max(min(20, int(math.ceil((float(r) / 80.0) * 20))), 1)
        sector = max(min(60, int(math.ceil((float(th) / 360.0) * 60))), 1)
        want[ring - 1, sector - 1] = max(want[ring - 1, sector - 1], float(zz))
    want[want == -1000.0] = 0
    got = oracle_mod.sc_make(p)
    assert np.array_equal(got, want)


# ---------------------------------------------------------------------------------------------- ikd-Tree Add_Points
def _add_points_python(existing, add, ds):
    """ikd_Tree.cpp:570-640 with downsample_on: per new point, box = floor(p / ds) * ds, the stored points inside it, winner =
    nearest to the box centre (strict <, the new point seeds the minimum); if more than one stored point is in the box or
    the new point wins, the box is emptied and the winner stored."""
    f32 = np.float32
    pts = [tuple(map(f32, r)) for r in existing]

    def d2(a, b):
        return (a[0] - b[0]) * (a[0] - b[0]) + (a[1] - b[1]) * (a[1] - b[1]) + (a[2] - b[2]) * (a[2] - b[2])

    for r in add:
        p = tuple(map(f32, r))
        mn = [f32(math.floor(float(p[k] / f32(ds)))) * f32(ds) for k in range(3)]
        mx = [mn[k] + f32(ds) for k in range(3)]
        mid = tuple(f32(float(mn[k]) + (float(mx[k]) - float(mn[k])) / 2.0) for k in range(3))
        inside = [j for j, s in enumerate(pts) if all(mn[k] <= s[k] < mx[k] for k in range(3))]
        best, win = d2(p, mid), p
        for j in inside:
            d = d2(pts[j], mid)
            if d < best:
                best, win = d, pts[j]
        same = all(abs(float(p[k]) - float(win[k])) < 1e-6 for k in range(3))
        if len(inside) > 1 or same:
            pts = [s for j, s in enumerate(pts) if j not in set(inside)] + [win]
        # else: exactly one stored point, and it is at least as close as the new one -> nothing changes
        elif len(inside) == 0:
            pts.append(p)
    return np.array(sorted(pts), np.float32)


def test_ikd_add_points_vs_python(oracle_mod):
    rng = np.random.default_rng(8)
    base = (rng.uniform(-2, 2, (300, 3)) * [1, 1, 0.1]).astype(np.float32)
    add = (rng.uniform(-2.4, 2.4, (400, 3)) * [1, 1, 0.1]).astype(np.float32)
    add[:60] = add[60:120] + rng.normal(0, 0.02, (60, 3)).astype(np.float32)
    got = oracle_mod.ikd_add_points(base, add, 0.4, True)
    want = _add_points_python(base, add, 0.4)
    g = np.array(sorted(map(tuple, got[:, :3])), np.float32)
    assert g.shape == want.shape and np.array_equal(g, want)


# ---------------------------------------------------------------------------------------------- feature labelling
def _label_python(p, curv, ring_start, ring_end):
    """scanRegistration.cpp:427-577 in plain Python: per ring, six segments in order; per segment a sort by curvature (ties
    by index: the declared rule where std::sort leaves it open), <= 2 sharp + <= 20 less-sharp picks from the top, <= 4
    flat picks from the bottom (the fourth without suppression), +-5 neighbour suppression stopped by a gap^2 > 0.05."""
    f32 = np.float32
    n = len(p)
    label = np.zeros(n, np.int32)
    picked = np.zeros(n, np.uint8)
    sharp, lsharp, flat = [], [], []

    def gap2(a, b):
        d = p[a] - p[b]
        return (d[0] * d[0] + d[1] * d[1]) + d[2] * d[2]

    def suppress(ind):
        picked[ind] = 1
        for l in range(1, 6):
            if float(gap2(ind + l, ind + l - 1)) > 0.05:
                break
            picked[ind + l] = 1
        for l in range(-1, -6, -1):
            if float(gap2(ind + l, ind + l + 1)) > 0.05:
                break
            picked[ind + l] = 1

    for s, e in zip(ring_start, ring_end):
        if e - s < 6:
            continue
        for j in range(6):
            sp, ep = s + (e - s) * j // 6, s + (e - s) * (j + 1) // 6 - 1
            order = sorted(range(sp, ep + 1), key=lambda i: (curv[i], i))
            largest = 0
            for ind in reversed(order):
                if picked[ind] == 0 and float(curv[ind]) > 0.1:
                    largest += 1
                    if largest <= 2:
                        label[ind] = 2
                        sharp.append(ind), lsharp.append(ind)
                    elif largest <= 20:
                        label[ind] = 1
                        lsharp.append(ind)
                    else:
                        break
                    suppress(ind)
            smallest = 0
            for ind in order:
                if picked[ind] == 0 and float(curv[ind]) < 0.1:
                    label[ind] = -1
                    flat.append(ind)
                    smallest += 1
                    if smallest >= 4:
                        break
                    suppress(ind)
    return label, sharp, lsharp, flat


def test_feature_labels_vs_python(oracle_mod, ilsm):
    c = ilsm.synth.config1(n_map=20_000)
    fe = oracle_mod.extract_features(c["cloud"])
    p = fe["cloud"][:, :3].astype(np.float32)
    label, sharp, lsharp, flat = _label_python(p, fe["curvature"], fe["ring_start"], fe["ring_end"])
    assert np.array_equal(label, fe["label"])
    assert sharp == fe["sharp_idx"].tolist() and lsharp == fe["less_sharp_idx"].tolist() and flat == fe["flat_idx"].tolist()
    assert len(sharp) > 100 and len(flat) > 500


def test_ring_assignment_vs_numpy(oracle_mod, ilsm):
    """scanRegistration.cpp:152-186, 277-374 (64-ring branch): min-range filter, angle = atan(z / sqrt(x^2 + y^2)) * 180 / pi
    (float), scanID = int((angle + 22.5) * 1.41 + 0.5) - 1, points outside [0, 63] dropped; the output is the rings
    concatenated, input order kept inside a ring; scanStartInd / scanEndInd carry the +5 / -6 margins (:387,393)."""
    c = ilsm.synth.config1(n_map=20_000)
    cloud = c["cloud"].astype(np.float32)
    fe = oracle_mod.extract_features(cloud)
    x, y, z = cloud[:, 0], cloud[:, 1], cloud[:, 2]
    with np.errstate(all="ignore"):
        ratio = (z / np.sqrt(x * x + y * y)).astype(np.float32)
        ang = (np.arctan(ratio.astype(np.float64)) * 180 / np.pi).astype(np.float32)
        sid = (ang.astype(np.float64) + 22.5) * 1.41 + 0.5
        sid = np.where(np.isfinite(sid), sid, -1.0).astype(np.int64) - 1          # int(): truncation toward zero
    keep = ((x * x + y * y) + z * z >= np.float32(0.3 * 0.3)) & (sid >= 0) & (sid <= 63)
    want_src = np.concatenate([np.where(keep & (sid == r))[0] for r in range(64)])
    assert np.array_equal(fe["src_index"], want_src)
    ids = fe["cloud"][:, 3].astype(np.int32)                                      # intensity = scanID + 0.1 * relTime
    assert np.array_equal(ids, sid[want_src])
    assert np.array_equal(fe["cloud"][:, :3], cloud[want_src, :3])
    for r in np.unique(ids):
        idx = np.where(ids == r)[0]
        assert fe["ring_start"][r] == idx[0] + 5 and fe["ring_end"][r] == idx[-1] + 1 - 6
    assert keep.sum() < len(cloud) * 0.6 and len(np.unique(ids)) > 30   # the OS0's +-45 deg beams beyond +-22.5 deg are dropped


# ---------------------------------------------------------------------------------------------- rolling cube map
class _CubeModel:
    """laserMapping.cpp:70-78, 330-606, 880-1002 with the cubes keyed by WORLD cube coordinates instead of a shifted array:
    rolling the 21 x 21 x 11 window then only means dropping the world cubes that leave it."""
    W, H, D = 21, 21, 11

    def __init__(self, line_res, plane_res):
        self.cen = [10, 10, 5]
        self.cubes = [{}, {}]  # corner, surf: world cube (wi, wj, wk) -> list of xyzi rows
        self.res = (line_res, plane_res)

    @staticmethod
    def _cube(v, cen):
        c = int((v + 25.0) / 50.0) + cen      # int(): truncation toward zero ...
        if v + 25.0 < 0:
            c -= 1                            # ... fixed up for negative coordinates (:333-338)
        return c

    def insert_world(self, corner, surf, centre):
        n = (self.W, self.H, self.D)
        ctr = [self._cube(float(centre[a]), self.cen[a]) for a in range(3)]
        for a in range(3):                    # :341-565: keep the centre at least 3 cubes away from the border
            while ctr[a] < 3:
                self.cen[a] += 1
                ctr[a] += 1
            while ctr[a] >= n[a] - 3:
                self.cen[a] -= 1
                ctr[a] -= 1
        for d in self.cubes:
            for w in [w for w in d if not all(0 <= w[a] + self.cen[a] < n[a] for a in range(3))]:
                del d[w]
        for which, pts in ((0, corner), (1, surf)):
            for r in np.asarray(pts, np.float32):   # :880-940: points outside the window are dropped
                idx = [self._cube(float(r[a]), self.cen[a]) for a in range(3)]
                if all(0 <= idx[a] < n[a] for a in range(3)):
                    row = np.zeros(4, np.float32)
                    row[:len(r)] = r[:4]
                    self.cubes[which].setdefault(tuple(idx[a] - self.cen[a] for a in range(3)), []).append(row)
        for i in range(ctr[0] - 2, ctr[0] + 3):     # :572-592 valid cubes, :987-1002 per-cube VoxelGrid
            for j in range(ctr[1] - 2, ctr[1] + 3):
                for k in range(ctr[2] - 1, ctr[2] + 2):
                    if 0 <= i < n[0] and 0 <= j < n[1] and 0 <= k < n[2]:
                        w = (i - self.cen[0], j - self.cen[1], k - self.cen[2])
                        for which in (0, 1):
                            if self.cubes[which].get(w):
                                self.cubes[which][w] = list(_voxelgrid_python(np.array(self.cubes[which][w]), self.res[which]))

    def all_points(self, which):
        rows = [r for v in self.cubes[which].values() for r in v]
        a = np.array(rows, np.float32).reshape(-1, 4)
        return a[np.lexsort((a[:, 2], a[:, 1], a[:, 0]))]


def test_rolling_cube_map_vs_world_keyed_model(oracle_mod):
    rng = np.random.default_rng(21)
    cm = oracle_mod.CubeMap(0.4, 0.8)
    model = _CubeModel(0.4, 0.8)
    # a path that drags the window through several rolls on every axis, negative coordinates included
    centres = [(0, 0, 0), (160, 0, 0), (420, -30, 0), (420, -380, 10), (90, -380, 140), (-260, 20, -40), (-262, 22, -41)]
    for ctr in centres:
        corner = (np.array(ctr) + rng.uniform(-70, 70, (300, 3)) * [1, 1, 0.3]).astype(np.float32)
        surf = (np.array(ctr) + rng.uniform(-70, 70, (900, 3)) * [1, 1, 0.3]).astype(np.float32)
        surf[:200] = surf[200:400] + rng.normal(0, 0.1, (200, 3)).astype(np.float32)   # voxel mates for the filter
        corner4, surf4 = np.zeros((300, 4), np.float32), np.zeros((900, 4), np.float32)
        corner4[:, :3], surf4[:, :3] = corner, surf
        corner4[:, 3], surf4[:, 3] = rng.integers(0, 64, 300), rng.integers(0, 64, 900)
        cm.insert_world(corner4, surf4, np.array(ctr, np.float64))
        model.insert_world(corner4, surf4, ctr)
        for which in (0, 1):
            got = [cm.cube(which, i, cap=2048) for i in range(21 * 21 * 11)]
            got = np.concatenate([g for g in got if len(g)]).reshape(-1, 4)
            got = got[np.lexsort((got[:, 2], got[:, 1], got[:, 0]))]
            want = model.all_points(which)
            assert got.shape == want.shape, (ctr, which, got.shape, want.shape)
            assert np.array_equal(got, want), (ctr, which)
    assert model.cen != [10, 10, 5]  # the window did roll


# ---------------------------------------------------------------------------------------------- odometry association
def _qrot(q, v):  # Eigen quaternion (x, y, z, w) applied to a vector, double
    u = np.array(q[:3])
    uv = 2.0 * np.cross(u, v)
    return v + q[3] * uv + np.cross(u, uv)


def _odom_associate_python(last_corner, last_surf, sharp, flat, qt):
    """laserOdometry.cpp:446-689 (DISTORTION 0): TransformToStart, 1-NN (d2 < 25), then the ring walk: edges take the nearest
    point on the rings (id, id + 2.5] / [id - 2.5, id); planes a second point on the same-or-nearer ring side and a third on
    the other rings.  Distances of the walk are float expressions (:478-483)."""
    f32 = np.float32
    q, t = np.array(qt[:4], np.float64), np.array(qt[4:], np.float64)
    out = []

    def sel(p):
        w = _qrot(q, p[:3].astype(np.float64)) + t
        return w.astype(np.float32)

    def nn(cloud, s):
        d = cloud[:, :3] - s
        d2 = (d[:, 0] * d[:, 0] + d[:, 1] * d[:, 1]) + d[:, 2] * d[:, 2]
        i = int(np.argmin(d2))
        return i, d2[i]

    def sq(cloud, j, s):
        d = cloud[j, :3] - s
        return float((d[0] * d[0] + d[1] * d[1]) + d[2] * d[2])

    ring_c, ring_s = last_corner[:, 3].astype(np.int32), last_surf[:, 3].astype(np.int32)
    for i, p in enumerate(sharp):
        s = sel(p)
        c, d2 = nn(last_corner, s)
        m2 = -1
        if d2 < f32(25.0):
            cid, best = int(ring_c[c]), 25.0
            for j in range(c + 1, len(last_corner)):
                if ring_c[j] <= cid:
                    continue
                if ring_c[j] > cid + 2.5:
                    break
                d = sq(last_corner, j, s)
                if d < best:
                    best, m2 = d, j
            for j in range(c - 1, -1, -1):
                if ring_c[j] >= cid:
                    continue
                if ring_c[j] < cid - 2.5:
                    break
                d = sq(last_corner, j, s)
                if d < best:
                    best, m2 = d, j
        if m2 >= 0:
            out.append((1, i, p[:3], last_corner[c, :3], last_corner[m2, :3]))
    for i, p in enumerate(flat):
        s = sel(p)
        c, d2 = nn(last_surf, s)
        m2 = m3 = -1
        if d2 < f32(25.0):
            cid, b2, b3 = int(ring_s[c]), 25.0, 25.0
            for j in range(c + 1, len(last_surf)):
                if ring_s[j] > cid + 2.5:
                    break
                d = sq(last_surf, j, s)
                if ring_s[j] <= cid and d < b2:
                    b2, m2 = d, j
                elif ring_s[j] > cid and d < b3:
                    b3, m3 = d, j
            for j in range(c - 1, -1, -1):
                if ring_s[j] < cid - 2.5:
                    break
                d = sq(last_surf, j, s)
                if ring_s[j] >= cid and d < b2:
                    b2, m2 = d, j
                elif ring_s[j] < cid and d < b3:
                    b3, m3 = d, j
        if m2 >= 0 and m3 >= 0:
            jj, ll, mm = (last_surf[k, :3].astype(np.float64) for k in (c, m2, m3))
            n = np.cross(jj - ll, jj - mm)      # LidarPlaneFactor: ljm_norm = (j - l) x (j - m), normalised (hpp:151-152)
            n /= np.linalg.norm(n)
            out.append((2, i, p[:3], n, np.array([-(jj @ n), 0, 0])))
    return out


def test_odometry_association_vs_python(oracle_mod, ilsm):
    import os, sys
    sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tools"))
    from sequence_bench import corridor_sequence
    clouds, _ = corridor_sequence(ilsm.synth, 2, 0x5EED0100, 40.0)
    f0, f1 = oracle_mod.extract_features(clouds[0]), oracle_mod.extract_features(clouds[1])
    last_corner = f0["cloud"][f0["less_sharp_idx"]]
    last_surf = f0["less_flat"]
    sharp, flat = f1["cloud"][f1["sharp_idx"]][:200], f1["cloud"][f1["flat_idx"]][:300]
    qt = np.array([0.001, -0.002, 0.004, 1.0, 0.18, 0.01, -0.005])
    qt[:4] /= np.linalg.norm(qt[:4])
    got = oracle_mod.odom_associate(last_corner, last_surf, sharp, flat, qt)
    want = _odom_associate_python(last_corner, last_surf, sharp, flat, qt)
    got = [g for g in got if g["type"] != 0]
    assert len(got) == len(want) and sum(w[0] == 1 for w in want) > 30 and sum(w[0] == 2 for w in want) > 100
    for g, w in zip(got, want):
        assert g["type"] == w[0] and g["src"] == w[1]
        assert np.array_equal(g["p"], w[2].astype(np.float64))
        if w[0] == 1:
            assert np.array_equal(g["a"], w[3].astype(np.float64)) and np.array_equal(g["b"], w[4].astype(np.float64))
        else:
            assert np.abs(g["a"] - w[3]).max() < 1e-12 and abs(g["b"][0] - w[4][0]) < 1e-10


# ---------------------------------------------------------------------------------------------- Ceres trust-region loop
def _quat_plus(x, d):
    """EigenQuaternionParameterization::Plus: [sin|d| / |d| * d, cos|d|] (x) x."""
    n = np.linalg.norm(d)
    if n == 0:
        return x.copy()
    dq = np.concatenate([np.sin(n) / n * d, [np.cos(n)]])
    a, b = dq, x
    return np.array([a[3] * b[0] + a[0] * b[3] + a[1] * b[2] - a[2] * b[1], a[3] * b[1] + a[1] * b[3] + a[2] * b[0] - a[0] * b[2],
                     a[3] * b[2] + a[2] * b[3] + a[0] * b[1] - a[1] * b[0], a[3] * b[3] - a[0] * b[0] - a[1] * b[1] - a[2] * b[2]])


def _ceres_lm_python(evaluate, x0, max_iter):
    """Ceres 1.14 TrustRegionMinimizer + LevenbergMarquardtStrategy with its defaults, on (cost, J^T J, J^T r) evaluations:
    Jacobi scaling fixed at iteration 0, D = sqrt(clamp(diag) / radius), rho test, radius update, the termination tests.
    The damped step is solved with numpy (the oracle uses an LDL^T it wrote itself)."""
    ftol, gtol, ptol = 1e-6, 1e-10, 1e-8
    x = np.array(x0, np.float64)
    cost, H, g = evaluate(x)
    initial = cost
    if not np.isfinite(cost):
        return x, 2, 0, 0, 0, initial, cost
    scale = 1.0 / (1.0 + np.sqrt(np.diag(H)))
    radius, decrease, it, ok_steps, bad_steps = 1e4, 2.0, 0, 0, 0
    diag = None
    reuse = False
    while True:
        if it >= max_iter:
            return x, 1, it, ok_steps, bad_steps, initial, cost          # NO_CONVERGENCE
        gmax = max(np.abs(g[3:]).max(), np.abs(x[:4] - _quat_plus(x[:4], -g[:3])).max())
        if gmax <= gtol:
            return x, 0, it, ok_steps, bad_steps, initial, cost
        if radius <= 1e-32:
            return x, 0, it, ok_steps, bad_steps, initial, cost
        it += 1
        Hs, gs = H * np.outer(scale, scale), g * scale
        if not reuse:
            diag = np.clip(np.diag(Hs), 1e-6, 1e32)
        reuse = True
        y = np.linalg.solve(Hs + np.diag(diag / radius), gs)
        model = 0.5 * (y @ gs + y @ (diag / radius * y))
        if not (np.isfinite(y).all() and model > 0):
            bad_steps += 1
            radius /= decrease
            decrease *= 2
            continue
        delta = -y * scale
        cand = np.concatenate([_quat_plus(x[:4], delta[:3]), x[4:] + delta[3:]])
        new_cost, nH, ng = evaluate(cand)
        if np.linalg.norm(cand - x) <= ptol * (np.linalg.norm(x) + ptol):
            return x, 0, it, ok_steps, bad_steps, initial, cost
        change = cost - new_cost
        if abs(change) <= ftol * cost:
            return x, 0, it, ok_steps, bad_steps, initial, cost
        rho = change / model
        if rho > 1e-3:
            x, cost, H, g = cand, new_cost, nH, ng
            radius = min(1e16, radius / max(1.0 / 3.0, 1.0 - (2.0 * rho - 1.0) ** 3))
            decrease, reuse = 2.0, False
            ok_steps += 1
        else:
            radius /= decrease
            decrease *= 2
            bad_steps += 1


@pytest.mark.parametrize("max_iter", [1, 4, 10, 30])
def test_trust_region_loop_vs_python(oracle_mod, cfg_small, max_iter):
    c = cfg_small
    qt0 = np.concatenate([c["q0"], c["t0"]])
    fac = oracle_mod.associate(c["map_corner"], c["map_surf"], c["corner"], c["surf"], qt0)
    x, s = oracle_mod.solve(fac, qt0, max_iter, 0.1)
    wx, term, it, ok, bad, ini, fin = _ceres_lm_python(lambda p: oracle_mod.evaluate(fac, p, 0.1), qt0, max_iter)
    assert (s.termination, s.iterations, s.num_successful, s.num_unsuccessful) == (term, it, ok, bad)
    assert abs(s.initial_cost - ini) <= 1e-12 * ini and abs(s.final_cost - fin) <= 1e-9 * fin
    assert np.abs(x - wx).max() < 1e-9
