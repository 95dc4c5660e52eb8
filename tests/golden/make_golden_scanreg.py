#!/usr/bin/env python
"""Generates tests/golden/scanreg_reference.npz: outputs of the REFERENCE's own LOAM front end
(/root/reference/src/scanRegistration.cpp:227-589, built into oracle/_ref/libref_scanreg.so by oracle/Makefile through
oracle/patches/scanreg_extract.py) on three seeded synthetic frames.  The frames are regenerated from their seeds by the
tests (numpy's generator is deterministic), so only the outputs are stored: sizes, SHA-256 of every output array, and the
per-label histogram.  Run in the build container (needs /root/reference):  python tests/golden/make_golden_scanreg.py"""
import functools
import hashlib
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
import oracle  # noqa: E402
import ilsm_b200 as ilsm  # noqa: E402


def frames():
    return _frames()


@functools.lru_cache(maxsize=1)
def _frames():
    return tuple(_gen_frames())


def _gen_frames():
    S = ilsm.synth
    scene = S.Scene()
    q0, t0 = S.default_pose()
    yield "open", S.make_frame(scene, q0, t0, seed=0x5EED0B01)[0]
    qc, tc = S.corridor_poses(3)[1]
    yield "corridor", S.make_frame(S.Scene(corridor=True, length=60.0), qc, tc, seed=0x5EED0B02)[0]
    yield "fov22", S.make_frame(scene, q0, t0, seed=0x5EED0B03, fov_deg=22.0)[0]


def digest(a):
    return hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()


KEYS = ("cloud", "curvature", "label", "sharp", "less_sharp", "flat", "less_flat_raw", "less_flat", "ring_start", "ring_end")
CLOUDS = ("cloud", "sharp", "less_sharp", "flat", "less_flat_raw", "less_flat")  # xyzi rows: xyz is also hashed alone


def as_reference_outputs(o):
    """An extract_features() dict (oracle or CUDA binding: index lists into the ring-ordered cloud) in the shape of the
    reference's outputs (feature clouds as xyzi rows).  The six segments of a ring tile [scanStartInd, scanEndInd - 1]
    (sp / ep of scanRegistration.cpp:440-441); curvature / label exist for i in [5, n - 5) only (:397-412)."""
    idx = [np.arange(s, e)[o["label"][s:e] <= 0] for s, e in zip(o["ring_start"], o["ring_end"]) if e - s >= 6]
    lf_raw = o["cloud"][np.concatenate(idx).astype(np.int64)] if idx else np.zeros((0, 4), np.float32)
    return {"cloud": o["cloud"], "curvature": o["curvature"][5:-5], "label": o["label"][5:-5].astype(np.int32),
            "sharp": o["cloud"][o["sharp_idx"]], "less_sharp": o["cloud"][o["less_sharp_idx"]], "flat": o["cloud"][o["flat_idx"]],
            "less_flat_raw": lf_raw, "less_flat": o["less_flat"], "ring_start": np.asarray(o["ring_start"], np.int32),
            "ring_end": np.asarray(o["ring_end"], np.int32)}

if __name__ == "__main__":
    out = {}
    for name, cloud in frames():
        r = oracle.ref_extract_features(cloud)
        # the reference computes curvature / label for i in [5, n - 5) only (:397-412); its global work arrays keep
        # whatever an earlier frame left outside that range
        r["curvature"], r["label"] = r["curvature"][5:-5], r["label"][5:-5]
        out[name + "/input_sha256"] = digest(cloud)
        for k in KEYS:
            out[f"{name}/{k}/shape"] = np.array(r[k].shape, np.int64)
            out[f"{name}/{k}/sha256"] = digest(r[k])
            if k in CLOUDS:
                out[f"{name}/{k}/xyz_sha256"] = digest(r[k][:, :3])
        out[name + "/label_hist"] = np.array([(r["label"] == v).sum() for v in (-1, 0, 1, 2)], np.int64)
        print(name, {k: r[k].shape for k in KEYS})
    np.savez(os.path.join(os.path.dirname(os.path.abspath(__file__)), "scanreg_reference.npz"), **out)
