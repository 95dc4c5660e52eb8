#!/usr/bin/env python
"""Configs 2 and 4 (BASELINE.json configs[1], configs[3]): the full odometry + mapping loop on a synthetic OS0-64
corridor sequence, one independent sequence per GPU.

  python tools/sequence_bench.py --frames 200 [--oracle-frames 50]
  python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port P \
         tools/sequence_bench.py --gpus N --frames 200

Frames are ray-cast on the host beforehand (not timed) into pinned memory.  Timed region per frame = ilsm_slam_frame:
H2D of the organised 64x1024 frame, front end, odometry, mapping, D2H of both poses (wall clock around the blocking
call; the call synchronises).  Rank 0 prints one JSON line: aggregate frames/s (max time over ranks), ATE vs ground
truth, and -- on the first --oracle-frames frames -- the CPU oracle's frames/s and the per-frame pose difference.
"""
from __future__ import annotations

import argparse
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def corridor_sequence(S, n_frames, seed0, length):
    scene = S.Scene(corridor=True, length=length)
    clouds, poses = [], []
    for k in range(n_frames):
        yaw = np.deg2rad(5.0) * np.sin(2 * np.pi * k / 50.0)
        q = S.quat_from_rotvec([0.0, 0.0, yaw])
        t = np.array([2.0 + 0.2 * k, 0.1 * np.sin(k / 15.0), 1.2])
        cloud, _ = S.make_frame(scene, q, t, seed=seed0 + k)
        clouds.append(cloud)
        poses.append((q, t))
    return clouds, poses


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--frames", type=int, default=200)
    ap.add_argument("--oracle-frames", type=int, default=50)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--pipelined", action="store_true",
                    help="laserMapping as its own stage (ilsm_slam_create_async): frame k's mapping overlaps frame k+1's front "
                         "end + odometry; the mapped pose of a frame comes back one call later")
    ap.add_argument("--sequences-per-gpu", type=int, default=1,
                    help="independent sequences replayed concurrently on each GPU (one context / stream / host thread each)")
    ap.add_argument("--mapping", default="laserMapping", choices=["laserMapping", "mapOptimization"],
                    help="mapping stage: the rolling cube map of laserMapping.cpp or the ground map of mapOptimization.cpp "
                         "(the node spot.launch starts)")
    args = ap.parse_args()
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))

    import torch
    import ilsm_b200 as ilsm

    S = ilsm.synth
    dist = None
    torch.cuda.set_device(local)
    dev = torch.device(f"cuda:{local}")
    if world > 1:
        import torch.distributed as dist_
        dist = dist_
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=dev)
    ilsm._build.build()
    F = args.frames
    clouds, poses = corridor_sequence(S, F, 0x5EED0100 + 4096 * rank, length=0.2 * F + 30.0)
    pinned = [torch.from_numpy(c).pin_memory() for c in clouds]
    views = [p.numpy() for p in pinned]
    ctx = ilsm.Context(local)

    def run(n, keep=False, slam=None):
        slam = slam or ilsm.Slam(ctx, 0.4, 0.8, 0.3, 8192, mapping=args.mapping, pipelined=args.pipelined)
        out, times = [], []
        for k in range(n):
            t0 = time.perf_counter()
            if args.pipelined:
                qo, to, qm, tm, st = slam.frame_async(views[k])
                if keep and k > 0:
                    out[-1] = out[-1][:2] + (qm, tm)  # the mapped pose belongs to the previous frame
                qm = tm = None
            else:
                qo, to, qm, tm, st = slam.frame(views[k])
            times.append(time.perf_counter() - t0)
            if keep:
                out.append((qo, to, qm, tm))
        if args.pipelined and n:
            t0 = time.perf_counter()
            qm, tm, st = slam.flush()
            times[-1] += time.perf_counter() - t0
            if keep:
                out[-1] = out[-1][:2] + (qm, tm)
        return out, times, slam

    run(min(args.warmup, F))[2].close()
    slam_timed = ilsm.Slam(ctx, 0.4, 0.8, 0.3, 8192, mapping=args.mapping, pipelined=args.pipelined)  # the 2 x 4851 x 8192-point cube slabs are allocated outside the timed region
    ctx.sync()
    if dist is not None:
        dist.barrier()
    torch.cuda.synchronize()
    l0 = ilsm.launch_count()
    t0 = time.perf_counter()
    slam_timed.host_phases()
    est, times, _ = run(F, keep=True, slam=slam_timed)
    phases = slam_timed.host_phases()
    torch.cuda.synchronize()
    wall = time.perf_counter() - t0
    slam_timed.close()  # freeing the 2.5 GB of cube slabs is not part of the per-frame loop
    launches = ilsm.launch_count() - l0
    SPG = max(1, args.sequences_per_gpu)
    if SPG > 1:
        # the per-frame chain is latency-bound (42 dependent launches, 3 host syncs): independent sequences interleave on
        # the SMs.  Every sequence gets its own context (stream + scratch) and host thread; ctypes releases the GIL.
        import threading
        ctxs = [ilsm.Context(local) for _ in range(SPG)]
        for c in ctxs:
            w = ilsm.Slam(c, 0.4, 0.8, 0.3, 8192, mapping=args.mapping)
            for k in range(min(3, F)):
                w.frame(views[k])
            w.close()
        slams = [ilsm.Slam(c, 0.4, 0.8, 0.3, 8192, mapping=args.mapping) for c in ctxs]
        finals = [None] * SPG

        def replay(i):
            for k in range(F):
                finals[i] = slams[i].frame(views[k])[3]

        th = [threading.Thread(target=replay, args=(i,)) for i in range(SPG)]
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        for t in th:
            t.start()
        for t in th:
            t.join()
        torch.cuda.synchronize()
        wall = time.perf_counter() - t0
        assert all(np.array_equal(f, est[-1][3]) for f in finals), "concurrent replay differs from the single-stream replay"
        for sl, c in zip(slams, ctxs):
            sl.close(), c.close()
    tt = torch.tensor([wall], dtype=torch.float64, device=dev)
    if dist is not None:
        dist.all_reduce(tt, op=dist.ReduceOp.MAX)
    wall_max = float(tt.cpu())

    if rank == 0:
        q0, t0p = poses[0]
        R0 = S.quat_to_mat(q0)
        err = []
        for k in range(F):
            tr = R0.T @ (poses[k][1] - t0p)
            err.append(np.linalg.norm(est[k][3] - tr))
        line = {"metric": "full odometry+mapping loop, frames/s (synthetic OS0-64 corridor)", "value": world * SPG * F / wall_max,
                "unit": "frames/s", "n_gpus": world, "sequences_per_gpu": SPG, "pipelined": bool(args.pipelined), "mapping": args.mapping, "frames_per_sequence": F, "ms_per_frame": 1e3 * wall_max / (F * SPG),
                "ms_per_frame_median": 1e3 * float(np.median(times)), "scaling": "weak",
                "h2d_bytes_per_frame": int(clouds[0].nbytes), "d2h_bytes_per_frame": 2 * 56 + 400,
                "gpu_launches_per_frame": launches / F,
                "host_phase_us_per_frame": dict(zip(["upload+fe_launch", "wait_prev_mapping", "wait_front_end", "odometry_launch",
                                                     "wait_odometry", "trees+mapping", "mapping_thread_launch", "mapping_thread_wait"],
                                                    [round(1e6 * x / F, 1) for x in phases[:8]])), "ate_rmse_m": float(np.sqrt(np.mean(np.square(err)))),
                "final_position_error_m": float(err[-1]),
                "slowest_frames_ms": [[int(i), round(1e3 * times[i], 3)] for i in np.argsort(times)[::-1][:6]],
                "timing": "host wall clock around ilsm_slam_frame (blocking), H2D of the frame and D2H of the poses inside"}
        no = min(args.oracle_frames, F)
        if no > 0 and world == 1:
            import oracle
            osl = oracle.Slam(0.4, 0.8, 0.3, mapping=args.mapping)
            dts, drs = [], []
            t1 = time.perf_counter()
            ores = [osl.frame(clouds[k]) for k in range(no)]
            cpu_wall = time.perf_counter() - t1
            for k in range(no):
                wmap = ores[k][1]
                dts.append(float(np.linalg.norm(est[k][3] - wmap[4:])))
                drs.append(float(S.quat_angle(est[k][2], wmap[:4])))
            line["cpu_baseline"] = {"value": no / cpu_wall, "unit": "frames/s", "cores": 1, "kind": "port",
                                    "sample": f"first {no} frames of the same sequence through the chained CPU oracle"}
            line["max_pose_diff_vs_oracle"] = {"m": max(dts), "rad": max(drs), "frames": no}
        print(json.dumps(line), flush=True)
    ctx.close()
    if dist is not None:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
