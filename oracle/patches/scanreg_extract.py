#!/usr/bin/env python
"""Cuts the numeric body of the reference's src/scanRegistration.cpp out of its ROS node so that it compiles without
ROS / PCL (TEST INFRASTRUCTURE for oracle/_ref/libref_scanreg.so).

laserCloudHandler (:189-669) is one function: image handling and ORB tracking first, then -- from `TicToc t_prepare`
(:227) to the per-ring VoxelGrid (:589) -- the LOAM front end this repository restates (removeClosedPointCloud, ring and
time tagging, curvature, the six-segment sort and sharp / less-sharp / flat / less-flat picking), then ROS publishing.
The file does not compile as it stands: the fork's annotation pass dropped the closing brace of the six-segment loop
(`for (j < 6)`) before :580, the closing brace of the `for (i < N_SCANS)` loop after :589 and the one of the function
("Read this first").  The repairs restore upstream A-LOAM's structure -- `sp` / `ep` are declared inside the
segment loop, so the less-flat collection (:567-577) belongs to it, and the per-ring VoxelGrid follows it once per ring.  This script writes three fragments, addressed by
line number of the pinned file (its SHA-256 is checked), which oracle/ref_scanreg.cpp includes between its own stand-ins
for pcl::PointCloud / pcl::VoxelGrid:

  globals.inc   :89 scanPeriod, :101 N_SCANS, :104-113 the four work arrays, :122 comp, :149 MINIMUM_RANGE
  remove.inc    :152-186 removeClosedPointCloud
  body.inc      :227-589 the front end, with three edits:
                  :235  `pcl::fromROSMsg(*laserCloudMsg, laserCloudIn);` -> `ref_fill(laserCloudIn);` (the wrapper's
                        copy of the caller's points: there is no ROS message here)
                  before :580  one `}` closes the `for (int j = 0; j < 6; j++)` loop the fork left open
                  after :589   one `}` closes the `for (int i = 0; i < N_SCANS; i++)` loop the fork left open

usage: scanreg_extract.py <reference scanRegistration.cpp> <output directory>   (a temporary build directory: the
fragments are never stored in this repository -- only oracle/_ref/libref_scanreg.so is kept, git-ignored)
"""
import hashlib
import os
import sys

PINNED_SHA256 = "7524657f370ff2696bdf689ec148465581bf62de0375767f78f9a8342a5b156c"
GLOBALS = [89, 101, 104, 107, 110, 113, 122, 149]
REMOVE = (152, 186)
BODY = (227, 589)
FROM_ROS_MSG = 235
CLOSE_SEGMENT_LOOP_BEFORE = 580


def main(src, out_dir):
    raw = open(src, "rb").read()
    got = hashlib.sha256(raw).hexdigest()
    if got != PINNED_SHA256:
        sys.exit(f"scanreg_extract: {src} is not the pinned file (sha256 {got}); the line-addressed cuts do not apply")
    lines = raw.decode("utf-8").split("\n")
    assert "pcl::fromROSMsg" in lines[FROM_ROS_MSG - 1]
    os.makedirs(out_dir, exist_ok=True)
    with open(os.path.join(out_dir, "globals.inc"), "w", encoding="utf-8") as f:
        f.write("\n".join(lines[n - 1] for n in GLOBALS) + "\n")
    with open(os.path.join(out_dir, "remove.inc"), "w", encoding="utf-8") as f:
        f.write("\n".join(lines[REMOVE[0] - 1:REMOVE[1]]) + "\n")
    body = lines[BODY[0] - 1:BODY[1]]
    assert "surfPointsLessFlatScanDS;" in lines[CLOSE_SEGMENT_LOOP_BEFORE - 1]
    body[FROM_ROS_MSG - BODY[0]] = "    ref_fill(laserCloudIn);"
    body.insert(CLOSE_SEGMENT_LOOP_BEFORE - BODY[0], "    }")
    body.append("}")
    with open(os.path.join(out_dir, "body.inc"), "w", encoding="utf-8") as f:
        f.write("\n".join(body) + "\n")


if __name__ == "__main__":
    main(sys.argv[1], sys.argv[2])
