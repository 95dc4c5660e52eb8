#!/usr/bin/env python
"""Cuts the scan-to-scan association of the reference's src/laserOdometry.cpp out of its ROS node so that it compiles
without ROS / PCL / Ceres / Eigen (TEST INFRASTRUCTURE for oracle/_ref/libref_laserodom.so).

The association lives inside main()'s spin loop: `for (opti_counter < 2 && use_aloam)` (:417-713) transforms every sharp
/ flat point to the start of the sweep (TransformToStart :147-172), finds its closest point in the previous frame's
less-sharp / less-flat cloud and walks the neighbouring rings for the second / third point (:446-689), hands the factors
to a ceres::Problem and calls ceres::Solve.  This script writes three fragments, addressed by line number of the pinned
file (its SHA-256 is checked), which oracle/ref_laserodom.cpp includes between stand-ins (pcl::PointCloud, a brute-force
pcl::KdTreeFLANN, a ceres::Problem that RECORDS the residual blocks it is given, a ceres::Solve that does nothing):

  globals.inc    :82 DISTORTION, :85 the two correspondence counters, :88-90 SCAN_PERIOD / DISTANCE_SQ_THRESHOLD /
                 NEARBY_SCAN, :105-106 the two k-d trees, :111-118 the feature clouds and the "last" clouds,
                 :130-131 para_q / para_t, :134-135 their Eigen::Map views
  transform.inc  :147-172 TransformToStart
  body.inc       :417-713 the two-pass association + solve loop, unmodified

usage: laserodom_extract.py <reference laserOdometry.cpp> <output directory>   (a temporary build directory: the
fragments are never stored in this repository -- only oracle/_ref/libref_laserodom.so is kept, git-ignored)
"""
import hashlib
import os
import sys

PINNED_SHA256 = "49a8c3fd4e39884415efdd80d38eda4ca30944dd6a13aa17c37cd30707c38612"
GLOBALS = [82, 85, 88, 89, 90, 105, 106, 111, 112, 113, 114, 117, 118, 130, 131, 134, 135]
TRANSFORM = (147, 172)
BODY = (417, 713)


def main(src, out_dir):
    raw = open(src, "rb").read()
    got = hashlib.sha256(raw).hexdigest()
    if got != PINNED_SHA256:
        sys.exit(f"laserodom_extract: {src} is not the pinned file (sha256 {got}); the line-addressed cuts do not apply")
    lines = raw.decode("utf-8").split("\n")
    assert "for (size_t opti_counter" in lines[BODY[0] - 1] and "void TransformToStart" in lines[TRANSFORM[0] - 1]
    os.makedirs(out_dir, exist_ok=True)
    for name, text in (("globals.inc", "\n".join(lines[n - 1] for n in GLOBALS)),
                       ("transform.inc", "\n".join(lines[TRANSFORM[0] - 1:TRANSFORM[1]])),
                       ("body.inc", "\n".join(lines[BODY[0] - 1:BODY[1]]))):
        with open(os.path.join(out_dir, name), "w", encoding="utf-8") as f:
            f.write(text + "\n")


if __name__ == "__main__":
    main(sys.argv[1], sys.argv[2])
