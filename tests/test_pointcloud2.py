"""sensor_msgs/PointCloud2 wire-format adapter: the message blob is uploaded as it is and unpacked by a
kernel (pcl::fromROSMsg's field map, scanRegistration.cpp:235).  CPU: the numpy restatement round-trips; GPU: bit-exact
against it for the Ouster layout, packed, unaligned and integer-intensity layouts, and the full loop fed with blobs gives
the same poses as fed with PCL-layout points."""
import ctypes

import numpy as np
import pytest

LAYOUTS = [  # point_step, x, y, z, intensity offset, intensity datatype
    (48, 0, 4, 8, 16, 7),    # Ouster driver (x y z pad intensity t reflectivity ring ambient range)
    (16, 0, 4, 8, 12, 7),    # packed xyzi
    (32, 0, 4, 8, 16, 7),    # pcl::PointXYZI as toROSMsg writes it
    (26, 2, 6, 10, 14, 4),   # unaligned fields, UINT16 intensity
    (20, 4, 8, 12, 0, 2),    # UINT8 intensity in front
    (24, 0, 4, 8, 16, 8),    # FLOAT64 intensity
    (24, 0, 4, 8, 12, 6),    # UINT32 intensity
    (12, 0, 4, 8, -1, 7),    # no intensity field
]


def _points(n, seed=0):
    rng = np.random.default_rng(seed)
    a = rng.normal(0, 20, (n, 4)).astype(np.float32)
    a[:, 3] = np.floor(rng.uniform(0, 250, n)).astype(np.float32)  # exactly representable in every intensity type
    a[::17, :3] = 0.0      # no-return rays
    a[5, 0] = np.nan       # is_dense = false clouds carry NaNs
    return a


@pytest.mark.parametrize("lay", LAYOUTS)
def test_oracle_pc2_round_trip(oracle_mod, lay):
    a = _points(1000)
    blob = oracle_mod.pc2_pack(a, *lay)
    assert blob.size == 1000 * lay[0]
    got = oracle_mod.pc2_unpack(blob, *lay)
    want = a.copy()
    if lay[4] < 0:
        want[:, 3] = 0
    assert np.array_equal(got.view(np.uint32), want.view(np.uint32))  # bit pattern, NaN included


def test_pc2_layout_struct_and_argument_checks(ilsm):
    lib = ilsm.load_library()
    lay = ilsm.pc2_layout_ouster()
    assert (lay.point_step, lay.off_x, lay.off_y, lay.off_z, lay.off_intensity, lay.intensity_datatype) == (48, 0, 4, 8, 16, 7)
    assert ctypes.sizeof(ilsm.Pc2Layout) == 32
    assert lib.ilsm_pc2_unpack(None, None, 0, ctypes.byref(lay), None) < 0  # null context: rejected before any CUDA call
    assert lib.ilsm_slam_frame_pc2(None, None, 0, ctypes.byref(lay), 1, None, None, None, None, None) < 0


def _layout(ilsm, lay):
    L = ilsm.Pc2Layout()
    L.point_step, L.off_x, L.off_y, L.off_z, L.off_intensity, L.intensity_datatype = lay
    return L


@pytest.mark.gpu
@pytest.mark.parametrize("lay", LAYOUTS)
def test_gpu_pc2_unpack_bit_exact(ctx, ilsm, oracle_mod, lay):
    a = _points(65536 + 37, seed=3)
    blob = oracle_mod.pc2_pack(a, *lay)
    got = ctx.pc2_unpack(blob, _layout(ilsm, lay))
    want = oracle_mod.pc2_unpack(blob, *lay)
    assert np.array_equal(got.view(np.uint32), want.view(np.uint32))


@pytest.mark.gpu
def test_gpu_pc2_rejects_bad_layouts_and_handles_empty(ctx, ilsm):
    assert ctx.pc2_unpack(np.zeros(0, np.uint8), ilsm.pc2_layout_ouster()).shape == (0, 4)
    for bad in ((8, 0, 4, 8, -1, 7), (16, 0, 4, 14, -1, 7), (16, 0, 4, 8, 12, 5), (16, 0, 4, 8, 13, 7)):
        with pytest.raises(ilsm.IlsmError):
            ctx.pc2_unpack(np.zeros(160, np.uint8), _layout(ilsm, bad))
    L = ilsm.pc2_layout_ouster()
    L.is_bigendian = 1
    with pytest.raises(ilsm.IlsmError):
        ctx.pc2_unpack(np.zeros(96, np.uint8), L)


@pytest.mark.gpu
def test_gpu_full_loop_from_message_blobs(ctx, ilsm, oracle_mod):
    """ilsm_slam_frame_pc2 (Ouster blobs, 48-byte points) == ilsm_slam_frame (packed points), pose for pose."""
    import sys, os
    sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tools"))
    from sequence_bench import corridor_sequence
    clouds, _ = corridor_sequence(ilsm.synth, 6, 0x5EED0100, 40.0)
    lay = LAYOUTS[0]
    a, b = ilsm.Slam(ctx, 0.4, 0.8, 0.3, 4096), ilsm.Slam(ctx, 0.4, 0.8, 0.3, 4096)
    for c in clouds:
        ra = a.frame(c)  # packed xyzi, 16-byte points
        rb = b.frame_pc2(oracle_mod.pc2_pack(c, *lay), _layout(ilsm, lay))
        for u, v in zip(ra[:4], rb[:4]):
            assert np.array_equal(u, v)
        assert ra[4].n_less_flat == rb[4].n_less_flat and ra[4].n_sharp == rb[4].n_sharp
    a.close(), b.close()


@pytest.mark.gpu
@pytest.mark.parametrize("lay", [l for l in LAYOUTS if l[5] == 7 and l[0] % 4 == 0 and l[1] % 4 == 0])
def test_gpu_pc2_pack_round_trip(ctx, ilsm, oracle_mod, lay):
    """ilsm_pc2_pack (pcl::toROSMsg of the published clouds): the blob unpacks to the same points (GPU and numpy
    restatement), the fields sit where the layout says and every other byte is zero; the default layout is PCL's
    PointXYZI (point_step 32, intensity at 16)."""
    a = _points(3001, seed=3)
    L = ilsm.Pc2Layout(lay[0], lay[1], lay[2], lay[3], lay[4], lay[5], 0, 0)
    blob = ctx.pc2_pack(a, L)
    assert blob.size == len(a) * lay[0]
    want = a.copy()
    if lay[4] < 0:
        want[:, 3] = 0
    assert np.array_equal(oracle_mod.pc2_unpack(blob, *lay).view(np.uint32), want.view(np.uint32))
    assert np.array_equal(ctx.pc2_unpack(blob, L).view(np.uint32), want.view(np.uint32))
    rec = blob.reshape(len(a), lay[0])
    used = np.zeros(lay[0], bool)
    for off in (lay[1], lay[2], lay[3]) + ((lay[4],) if lay[4] >= 0 else ()):
        used[off:off + 4] = True
    assert not rec[:, ~used].any()
    d = ctx.pc2_pack(a)  # default: pcl::PointXYZI
    assert d.size == len(a) * 32 and np.array_equal(d.reshape(-1, 32)[:, 16:20].view(np.float32)[:, 0].view(np.uint32), a[:, 3].view(np.uint32))
    assert ctx.pc2_pack(np.zeros((0, 4), np.float32)).size == 0
    with pytest.raises(ilsm.IlsmError):
        ctx.pc2_pack(a, ilsm.Pc2Layout(26, 2, 6, 10, 14, 4, 0, 0))  # unaligned / integer intensity cannot be written
