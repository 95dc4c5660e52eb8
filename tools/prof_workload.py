#!/usr/bin/env python
"""One pass over every hot kernel at representative sizes, bracketed by cudaProfilerStart/Stop, for
  ncu --set full --clock-control none --import-source on --profile-from-start off -o gpurun_out/prof python tools/prof_workload.py
Sections (select with --only): reg (config-1 registration: builds, associate, solve), knn (N = 2M, Q = 65536),
jtj (1M factors), sc (20k-keyframe shard), fe (projection + feature extraction), slam (one frame of the full loop after
20 frames of a corridor sequence), slamlong (one frame after 600 frames of the bench's sequence)."""
from __future__ import annotations

import argparse
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--only", default="reg,knn,jtj,sc,fe")
    args = ap.parse_args()
    only = set(args.only.split(","))
    import torch
    import ilsm_b200 as ilsm

    dev = torch.device("cuda:0")
    ctx = ilsm.Context(0)
    S = ilsm.synth
    opts = ilsm.default_opts()

    def pad4(a):
        o = np.zeros((len(a), 4), np.float32)
        o[:, :3] = a[:, :3]
        return torch.from_numpy(o).to(dev)

    work = []
    if "reg" in only:
        c = S.config1(n_map=100_000)
        d_mc, d_ms, d_c, d_s = pad4(c["map_corner"]), pad4(c["map_surf"]), pad4(c["corner"]), pad4(c["surf"])
        pose0 = torch.from_numpy(np.concatenate([c["q0"], c["t0"]])).to(dev)
        pose = pose0.clone()
        mc, ms = ctx.new_map(), ctx.new_map()

        def reg():
            pose.copy_(pose0)
            mc.build_pair_dev(d_mc.data_ptr(), len(d_mc), ms, d_ms.data_ptr(), len(d_ms), 16)
            ctx.register_dev(mc, ms, d_c.data_ptr(), len(d_c), d_s.data_ptr(), len(d_s), 16, pose.data_ptr(), opts)
        work.append(reg)
    if "knn" in only or "jtj" in only:
        c2 = S.config1(n_map=2_000_000)
        d_map = pad4(np.concatenate([c2["map_corner"], c2["map_surf"]]))
        gm = ctx.new_map().build_dev(d_map.data_ptr(), len(d_map), 16)
        R = S.quat_to_mat(c2["q_true"])
        world = (c2["cloud"][:, :3].astype(np.float64) @ R.T + c2["t_true"]).astype(np.float32)
        d_q = pad4(world)
        d_idx = torch.empty((65536, 5), dtype=torch.int32, device=dev)
        d_d2 = torch.empty((65536, 5), dtype=torch.float32, device=dev)
        if "knn" in only:
            work.append(lambda: gm.knn_dev(d_q.data_ptr(), 65536, 16, 5, 0.0, d_idx.data_ptr(), d_d2.data_ptr()))  # exact
        if "jtj" in only:
            d_mc2, d_ms2 = pad4(c2["map_corner"]), pad4(c2["map_surf"])
            mc2 = ctx.new_map().build_dev(d_mc2.data_ptr(), len(d_mc2), 16)
            ms2 = ctx.new_map().build_dev(d_ms2.data_ptr(), len(d_ms2), 16)
            Qb = 1 << 22
            sens = np.zeros((Qb, 4), np.float32)
            sens[:, :3] = c2["cloud"][np.arange(Qb) % 65536, :3]
            d_c2, d_s2 = torch.from_numpy(sens[:Qb // 8].copy()).to(dev), torch.from_numpy(sens[Qb // 8:].copy()).to(dev)
            pose_t = torch.from_numpy(np.concatenate([c2["q_true"], c2["t_true"]])).to(dev)
            out32 = torch.zeros(32, dtype=torch.float64, device=dev)

            def jtj():
                ctx.associate_dev(mc2, ms2, d_c2.data_ptr(), Qb // 8, d_s2.data_ptr(), Qb - Qb // 8, 16, pose_t.data_ptr(), opts)
                ctx.eval_normal_eq_dev(pose_t.data_ptr(), out32.data_ptr())
            work.append(jtj)
    if "sc" in only:
        db = S.sc_database(20_000)
        q, _, _ = S.sc_queries(db, 2)
        sc = ilsm.ScanContextDb(ctx)
        sc.add(db)
        d_sq = torch.from_numpy(q.reshape(2, 1200)).to(dev)
        pack = torch.zeros(160, dtype=torch.uint8, device=dev)
        work.append(lambda: sc.query_packed_dev(d_sq[0].data_ptr(), 10, 20_000, 0, pack.data_ptr()))
    if "sc8" in only:  # config 5 shape: one batch of 8 queries against a 100k-keyframe shard (prefilter + selection + exact rescoring)
        scb = ilsm.ScanContextDb(ctx)
        for a in range(0, 100_000, 20_000):
            scb.add(S.sc_database_range(a, a + 20_000, 100_000))
        q8, _, _ = S.sc_chunked_queries(100_000, 8)
        d_q8 = torch.from_numpy(q8.reshape(8, 1200)).to(dev)
        pack8 = torch.zeros(8 * 160, dtype=torch.uint8, device=dev)
        work.append(lambda: scb.query_topk_sharded_dev(d_q8.data_ptr(), 8, 10, 99_950, 0, pack8.data_ptr()))
    if "sc64" in only:  # the bench's batch: 64 queries per call against a 100k-keyframe shard
        scc = ilsm.ScanContextDb(ctx)
        for a in range(0, 100_000, 20_000):
            scc.add(S.sc_database_range(a, a + 20_000, 100_000))
        q64, _, _ = S.sc_chunked_queries(100_000, 64)
        d_q64 = torch.from_numpy(q64.reshape(64, 1200)).to(dev)
        pack64 = torch.zeros(64 * 160, dtype=torch.uint8, device=dev)
        work.append(lambda: scc.query_topk_sharded_dev(d_q64.data_ptr(), 64, 10, 99_950, 0, pack64.data_ptr()))
    if "fe" in only:
        c3 = S.config1(n_map=20_000)
        cloud = c3["cloud"]

        def fe():
            ctx.cloud_handler(cloud)
            ctx.extract_features(cloud)
        work.append(fe)

    if "slam" in only:
        sys.path.insert(0, os.path.join(ROOT, "tools"))
        from sequence_bench import corridor_sequence
        clouds, _ = corridor_sequence(S, 24, 0x5EED0100, 40.0)
        slam = ilsm.Slam(ctx, 0.4, 0.8, 0.3, 8192)
        for k in range(20):
            slam.frame(clouds[k])
        it = iter(clouds[20:])
        work.append(lambda: slam.frame(next(it)))

    if "slamlong" in only:  # one frame of the full loop deep into the bench's corridor sequence (cubes filled)
        nf = 604
        scene = S.Scene(corridor=True, length=0.2 * nf + 30.0)
        lclouds = S.make_frames_torch(scene, S.corridor_poses(nf), 0x5EED0200, dev)
        lviews = [lclouds[k].numpy() for k in range(nf)]
        lslam = ilsm.Slam(ctx, 0.4, 0.8, 0.3, 8192)
        for k in range(600):
            lslam.frame(lviews[k])
        lit = iter(lviews[600:])
        work.append(lambda: lslam.frame(next(lit)))

    for _ in range(3):  # warm-up (allocation, module load; the map builds alternate between two tables)
        for w in work:
            w()
    ctx.sync()
    torch.cuda.synchronize()
    torch.cuda.profiler.start()
    for w in work:
        w()
    ctx.sync()
    torch.cuda.synchronize()
    torch.cuda.profiler.stop()
    print("prof_workload done")


if __name__ == "__main__":
    main()
