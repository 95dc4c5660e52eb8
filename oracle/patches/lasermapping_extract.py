#!/usr/bin/env python
"""Cuts the map maintenance of the reference's src/laserMapping.cpp out of its ROS node so that it compiles without ROS /
PCL / Ceres / Eigen (TEST INFRASTRUCTURE for oracle/_ref/libref_lasermapping.so).

process() (:233-1166) is one loop body: queue handling, then transformAssociateToMap and the rolling 21x21x11 cube window
(:327-565), the 5x5x3 gather (:566-593), the stack VoxelGrids (:595-603), the guarded optimisation (:624-873), transformUpdate
(:875), the insertion of the stack into the cubes (:878-945) and the per-cube VoxelGrid of the valid cubes (:984-1002).  The
file does not compile as it stands: the fork's annotation pass duplicated the surf insertion loop together with a second
`TicToc t_filter;` declaration in the same scope (:946-983; SURVEY.md "fork defects").  This script writes three fragments,
addressed by line number of the pinned file (its SHA-256 is checked), which oracle/ref_lasermapping.cpp includes between
stand-ins (pcl::PointCloud, pcl::VoxelGrid = the oracle's restatement, Eigen = oracle/shims):

  globals.inc    :63-115 and :123-129 (window geometry, the cube arrays, the pose parameters and their Eigen::Map views,
                 the two transforms, the two VoxelGrid filters) -- without :117-121, the ROS message queues and their mutex
  transform.inc  :138-172 transformAssociateToMap, transformUpdate, pointAssociateToMap, pointAssociateTobeMapped
  body.inc       :327-623  window roll, gather, stack VoxelGrids
                 :875-945  transformUpdate, insertion                      (the optimisation block :624-874 is NOT included:
                 :984-1004 per-cube VoxelGrid of the valid cubes            both sides of the comparison keep the predicted pose)
                 i.e. the duplicated :946-983 is dropped, as upstream A-LOAM has it (one insertion, one t_filter)
  optimise.inc   :624-873  the guarded optimisation block on its own (two passes of 5-NN association, line / plane fits,
                 residual blocks, ceres::Solve), unmodified -- ref_lasermapping_associate runs it with a ceres::Problem
                 that records the blocks and a Solve that does nothing

usage: lasermapping_extract.py <reference laserMapping.cpp> <output directory>   (a temporary build directory: the
fragments are never stored in this repository -- only oracle/_ref/libref_lasermapping.so is kept, git-ignored)
"""
import hashlib
import os
import sys

PINNED_SHA256 = "0ad4912266f6afdcefa6ffdc815bbf5908cf7ac5ef85624f77ee9e8e39939c24"
GLOBALS = [(63, 115), (123, 129)]
TRANSFORM = (138, 172)
BODY = [(327, 623), (875, 945), (984, 1004)]
OPTIMISE = (624, 873)


def main(src, out_dir):
    raw = open(src, "rb").read()
    got = hashlib.sha256(raw).hexdigest()
    if got != PINNED_SHA256:
        sys.exit(f"lasermapping_extract: {src} is not the pinned file (sha256 {got}); the line-addressed cuts do not apply")
    lines = raw.decode("utf-8").split("\n")
    assert "transformAssociateToMap();" in lines[327 - 1] and "transformUpdate();" in lines[875 - 1] and "TicToc t_filter;" in lines[984 - 1]
    cut = lambda ranges: "\n".join("\n".join(lines[a - 1:b]) for a, b in ranges) + "\n"
    os.makedirs(out_dir, exist_ok=True)
    assert "if (laserCloudCornerFromMapNum > 10" in lines[OPTIMISE[0] - 1]
    for name, text in (("globals.inc", cut(GLOBALS)), ("transform.inc", cut([TRANSFORM])), ("body.inc", cut(BODY)),
                       ("optimise.inc", cut([OPTIMISE]))):
        with open(os.path.join(out_dir, name), "w", encoding="utf-8") as f:
            f.write(text)


if __name__ == "__main__":
    main(sys.argv[1], sys.argv[2])
