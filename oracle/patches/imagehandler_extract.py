#!/usr/bin/env python
"""Cuts the range / intensity image projection of the reference's src/image_handler.h_ouster out of its ROS + OpenCV class
(TEST INFRASTRUCTURE for oracle/_ref/libref_imagehandler.so).

ImageHandler::cloud_handler (:103-140) converts the ROS message, allocates three cv::Mat images and then runs one loop over
the organised cloud (:113-139): range = |p|, the 8-bit range image min(20 range, 255), the 8-bit intensity image
min(intensity, 255), and the cloud_track copy with points below 0.1 m zeroed.  This script writes that loop, addressed by line
number of the pinned file (its SHA-256 is checked); oracle/ref_imagehandler.cpp includes it between stand-ins for cv::Mat
(a byte matrix with at<uint8_t>(u, v)) and the two PCL point clouds.

  loop.inc   :113-139, unmodified

usage: imagehandler_extract.py <reference image_handler.h_ouster> <output directory>   (a temporary build directory: the
fragment is never stored in this repository -- only oracle/_ref/libref_imagehandler.so is kept, git-ignored)
"""
import hashlib
import os
import sys

PINNED_SHA256 = "84c0e05f12cc348ae5d6c331e30f099253ce4860b2ff4abf13f963d603d1cfab"
LOOP = (113, 139)


def main(src, out_dir):
    raw = open(src, "rb").read()
    got = hashlib.sha256(raw).hexdigest()
    if got != PINNED_SHA256:
        sys.exit(f"imagehandler_extract: {src} is not the pinned file (sha256 {got}); the line-addressed cut does not apply")
    lines = raw.decode("utf-8").split("\n")
    assert "for (int u = 0; u < IMAGE_HEIGHT; u++)" in lines[LOOP[0] - 1]
    os.makedirs(out_dir, exist_ok=True)
    with open(os.path.join(out_dir, "loop.inc"), "w", encoding="utf-8") as f:
        f.write("\n".join(lines[LOOP[0] - 1:LOOP[1]]) + "\n")


if __name__ == "__main__":
    main(sys.argv[1], sys.argv[2])
