#!/usr/bin/env python
"""Full loop driven from C++ (tools/cpp/slam_replay.cpp): writes a synthetic corridor sequence to a file, builds the C++
driver against libilsm_cuda.so, replays the file and prints its JSON line next to the Python path's frames/s and the
pose difference between the two (identical C ABI calls, so the poses must be bit-identical).

  python tools/cpp_replay.py [--frames 150] [--sequences 4]
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import tempfile
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tools"))
PKG = os.path.join(ROOT, "intensity_based_lidar_slam_for_me-_b200")


def build_driver(out_dir):
    exe = os.path.join(out_dir, "slam_replay")
    cmd = ["g++", "-std=c++14", "-O2", "-Wall", "-I", os.path.join(ROOT, "include"), os.path.join(ROOT, "tools", "cpp", "slam_replay.cpp"),
           "-o", exe, "-L", PKG, "-lilsm_cuda", "-pthread", "-Wl,-rpath," + PKG]
    subprocess.run(cmd, check=True)
    return exe


def run(frames=150, sequences=1, out_dir=None, pipelined=False):
    import ilsm_b200 as ilsm
    from sequence_bench import corridor_sequence
    ilsm._build.build()
    out_dir = out_dir or tempfile.mkdtemp(prefix="ilsm_replay_")
    clouds, _ = corridor_sequence(ilsm.synth, frames, 0x5EED0100, 0.2 * frames + 30.0)
    path = os.path.join(out_dir, "frames.bin")
    np.stack(clouds).astype(np.float32).tofile(path)
    exe = build_driver(out_dir)
    r = subprocess.run([exe, path, str(sequences), "1" if pipelined else "0"], capture_output=True, text=True, timeout=600)
    if r.returncode != 0:
        raise RuntimeError(r.stdout + r.stderr)
    cpp = json.loads(r.stdout.strip().splitlines()[-1])
    cpp["host_phases"] = r.stderr.strip().splitlines()[-1] if r.stderr.strip() else ""
    # the same replay through the Python binding
    ctx = ilsm.Context(0)
    slam = ilsm.Slam(ctx, 0.4, 0.8, 0.3, 8192)
    for k in range(5):
        slam.frame(clouds[k])
    slam.close()
    slam = ilsm.Slam(ctx, 0.4, 0.8, 0.3, 8192)
    t0 = time.perf_counter()
    for k in range(frames):
        qo, to, qm, tm, _ = slam.frame(clouds[k])
    wall = time.perf_counter() - t0
    slam.close(), ctx.close()
    os.remove(path)
    return {"cpp": cpp, "python_frames_per_s": frames / wall,
            "pose_identical": bool(np.array_equal(qm, cpp["q_map"]) and np.array_equal(tm, cpp["t_map"]) and
                                   np.array_equal(qo, cpp["q_odom"]) and np.array_equal(to, cpp["t_odom"]))}


if __name__ == "__main__":
    ap = argparse.ArgumentParser()
    ap.add_argument("--frames", type=int, default=150)
    ap.add_argument("--sequences", type=int, default=1)
    ap.add_argument("--pipelined", action="store_true")
    a = ap.parse_args()
    print(json.dumps(run(a.frames, a.sequences, pipelined=a.pipelined)))
