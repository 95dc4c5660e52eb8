#!/usr/bin/env python
"""Counts the SASS opcodes that prove which hardware mechanisms each kernel uses:
  cuobjdump -sass intensity_based_lidar_slam_for_me-_b200/libilsm_cuda.so | python tools/sass_mechanisms.py > profiles/r02_sass_mechanisms.txt"""
import collections
import re
import subprocess
import sys

KEEP = ("UBLKCP", "STAS", "SYNCS", "UCGABAR_ARV", "UCGABAR_WAIT", "HMMA", "LDSM", "REDUX", "CREDUX", "LDGSTS", "MATCH", "DMMA", "UTMALDG",
        "UTCHMMA", "LDTM")
cur = None
cnt = collections.defaultdict(collections.Counter)
for line in sys.stdin:
    m = re.search(r"Function : (\S+)", line)
    if m:
        cur = m.group(1)
        continue
    mm = re.search(r"/\*[0-9a-f]{4,}\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_]+(?:\.[A-Z0-9_\.]+)?)", line)
    if cur is None or not mm:
        continue
    op = mm.group(1)
    if op.split(".")[0] in KEEP:
        cnt[cur][op] += 1


def dem(n):
    try:
        return subprocess.run(["c++filt", n], capture_output=True, text=True).stdout.strip().split("(")[0]
    except Exception:
        return n


print("# SASS mechanisms per kernel of libilsm_cuda.so (sm_100a), round 2")
print("# cuobjdump -sass libilsm_cuda.so | python tools/sass_mechanisms.py -- opcodes counted per function:")
print("# bulk async copies (UBLKCP = cp.async.bulk, the 1-D TMA path), DSMEM pushes (STAS), transaction barriers (SYNCS), cluster")
print("# barrier (UCGABAR), tensor-core MMA (HMMA) + ldmatrix (LDSM), warp reductions (REDUX, CREDUX = redux.sync.min/max on sm_100), cp.async (LDGSTS), warp match (MATCH).")
print("# No tcgen05 (UTC*MMA / LDTM) and no tensor-map TMA (UTMALDG): nothing on this path is a large dense contraction; the one")
print("# contraction (the ScanContext prefilter, K = 20 rings / 64 sector keys) uses mma.sync m16n8k16 (HMMA.16816.F32).")
for k in sorted(cnt, key=dem):
    print(dem(k))
    for op, c in sorted(cnt[k].items()):
        print(f"    {op:44s} {c}")
