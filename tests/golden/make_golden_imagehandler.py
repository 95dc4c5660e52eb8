#!/usr/bin/env python
"""Generates tests/golden/imagehandler_reference.npz: outputs of the REFERENCE's own projection loop
(/root/reference/src/image_handler.h_ouster:113-139, built into oracle/_ref/libref_imagehandler.so by oracle/Makefile through
oracle/patches/imagehandler_extract.py) on a seeded organised 64 x 1024 frame with intensities on both sides of 255, ranges on
both sides of 12.75 m (the 8-bit range image saturates there) and points below the 0.1 m cut: SHA-256 of the three outputs.
Run in the build container (needs /root/reference):  python tests/golden/make_golden_imagehandler.py"""
import functools
import hashlib
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
import ilsm_b200 as ilsm  # noqa: E402


@functools.lru_cache(maxsize=1)
def frame():
    S = ilsm.synth
    cloud = S.make_frame(S.Scene(), *S.default_pose(), seed=0x5EED0F00)[0].copy()
    rng = np.random.default_rng(1)
    cloud[:, 3] = rng.uniform(0, 400, len(cloud)).astype(np.float32)
    cloud[:50, :3] *= 0.001
    return cloud


def digests(image_range, image_intensity, track):
    return [hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest() for a in (image_range, image_intensity, track)]


if __name__ == "__main__":
    import oracle
    r = oracle.ref_project(frame())
    print("range max", r[0].max(), "intensity max", r[1].max(), "zeroed track points", int((r[2][:, :3] == 0).all(1).sum()))
    np.savez(os.path.join(os.path.dirname(os.path.abspath(__file__)), "imagehandler_reference.npz"), sha256=np.array(digests(*r)),
             range_hist=np.bincount(r[0].ravel(), minlength=256), intensity_hist=np.bincount(r[1].ravel(), minlength=256))
