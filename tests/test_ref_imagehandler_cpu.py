"""The oracle's range / intensity image projection against the REFERENCE's own loop (src/image_handler.h_ouster:113-139 cut
out of ImageHandler::cloud_handler and compiled into oracle/_ref/libref_imagehandler.so): live when the library exists, and
against its committed outputs everywhere."""
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "tests", "golden"))
from make_golden_imagehandler import digests, frame  # noqa: E402

GOLD = np.load(os.path.join(ROOT, "tests", "golden", "imagehandler_reference.npz"))


def test_oracle_projection_equals_reference_golden(oracle_mod):
    r = oracle_mod.project(frame())
    assert np.array_equal(np.bincount(r[0].ravel(), minlength=256), GOLD["range_hist"])
    assert np.array_equal(np.bincount(r[1].ravel(), minlength=256), GOLD["intensity_hist"])
    assert digests(*r) == [str(s) for s in GOLD["sha256"]]


def test_oracle_projection_equals_reference_live(oracle_mod):
    if oracle_mod.ref_imagehandler() is None:
        pytest.skip("oracle/_ref/libref_imagehandler.so not built (needs the reference tree)")
    rng = np.random.default_rng(3)
    for k in range(3):
        c = frame().copy()
        c[:, :3] *= rng.uniform(0.02, 1.5, (len(c), 1)).astype(np.float32)
        c[:, 3] = rng.uniform(0, 600, len(c)).astype(np.float32)
        a, b = oracle_mod.ref_project(c), oracle_mod.project(c)
        assert all(np.array_equal(x, y) for x, y in zip(a, b)), k
