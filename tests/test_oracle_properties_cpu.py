"""Randomised properties of the oracle (hypothesis): k-d tree == brute force under the (d2, index) tie-break on clouds with
duplicates and lattice ties; VoxelGrid invariants; the down-sampled insertion keeps at most one point per touched box."""
import numpy as np
from hypothesis import given, settings, strategies as st


def _cloud(seed, n, lattice):
    rng = np.random.default_rng(seed)
    p = rng.normal(0, 2.0, (n, 3)).astype(np.float32)
    if lattice:
        p = np.round(p * 2) / 2          # many exact ties and duplicates
    return p.astype(np.float32)


@settings(max_examples=40, deadline=None)
@given(seed=st.integers(0, 2**31 - 1), n=st.integers(1, 400), nq=st.integers(1, 60), k=st.sampled_from([1, 5, 8]), lattice=st.booleans())
def test_kdtree_equals_brute_force(oracle_mod, seed, n, nq, k, lattice):
    m, q = _cloud(seed, n, lattice), _cloud(seed + 1, nq, lattice)
    bi, bd = oracle_mod.knn_brute(m, q, k)
    ti, td = oracle_mod.knn_kdtree(m, q, k)
    assert np.array_equal(bi, ti) and np.array_equal(bd, td)
    found = bi >= 0
    assert (found.sum(axis=1) == min(k, n)).all()
    d = np.where(found, bd, np.inf)
    assert (d[:, :-1] <= d[:, 1:]).all()                         # ascending distances (missing neighbours = +inf last)
    for r in range(nq):                                          # ties in ascending index
        for a in range(min(k, n) - 1):
            if bd[r, a] == bd[r, a + 1]:
                assert bi[r, a] < bi[r, a + 1]


@settings(max_examples=30, deadline=None)
@given(seed=st.integers(0, 2**31 - 1), n=st.integers(1, 600), leaf=st.sampled_from([0.2, 0.4, 0.8]))
def test_voxelgrid_invariants(oracle_mod, seed, n, leaf):
    rng = np.random.default_rng(seed)
    c = np.zeros((n, 4), np.float32)
    c[:, :3] = rng.normal(0, 1.5, (n, 3))
    c[:, 3] = rng.integers(0, 64, n)
    out = oracle_mod.voxelgrid(c, leaf)
    assert 1 <= len(out) <= n
    # every centroid lies inside the bounding box of the input, intensities inside the input range
    assert (out[:, :3] >= c[:, :3].min(0) - 1e-5).all() and (out[:, :3] <= c[:, :3].max(0) + 1e-5).all()
    assert out[:, 3].min() >= c[:, 3].min() - 1e-4 and out[:, 3].max() <= c[:, 3].max() + 1e-4
    # filtering the output again with the same leaf cannot increase the count, and a single point is a fixed point
    assert len(oracle_mod.voxelgrid(out, leaf)) <= len(out)
    one = c[:1]
    assert np.array_equal(oracle_mod.voxelgrid(one, leaf), one)
    # the point count is conserved: total mass = n (centroid * count summed over voxels equals the sum of the inputs)
    inv = np.float32(1.0) / np.float32(leaf)
    keys = np.floor(c[:, :3] * inv).astype(np.int64)
    assert len(out) == len(np.unique(keys, axis=0))


@settings(max_examples=25, deadline=None)
@given(seed=st.integers(0, 2**31 - 1), n_old=st.integers(0, 150), n_add=st.integers(1, 150))
def test_downsampled_insertion_invariant(oracle_mod, seed, n_old, n_add):
    rng = np.random.default_rng(seed)
    ds = 0.4
    old = rng.uniform(-1.5, 1.5, (n_old, 3)).astype(np.float32)
    add = rng.uniform(-1.5, 1.5, (n_add, 3)).astype(np.float32)
    out = oracle_mod.ikd_add_points(old, add, ds, True)[:, :3]
    box = lambda p: np.floor(p / np.float32(ds)).astype(np.int64)
    touched = {tuple(b) for b in box(add)}
    counts = {}
    for b in map(tuple, box(out)):
        counts[b] = counts.get(b, 0) + 1
    for b in touched:                      # every box a new point fell into ends with exactly one point
        assert counts.get(b, 0) == 1
    untouched_old = [tuple(b) for b in box(old) if tuple(b) not in touched] if n_old else []
    for b in set(untouched_old):           # untouched boxes keep all their (Build-seeded) points
        assert counts[b] == untouched_old.count(b)
    # appending without down-sampling keeps everything, in order
    app = oracle_mod.ikd_add_points(old, add, ds, False)[:, :3]
    assert np.array_equal(app, np.concatenate([old, add]))
