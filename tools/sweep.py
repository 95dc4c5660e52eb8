#!/usr/bin/env python
"""Config 3 (BASELINE.json configs[2]): dense local-map sweep N = 100k .. 2M points, Q = 4096 / 65536 queries;
the k-NN kernel alone, the fused associate (k-NN + fit) kernel alone and the J^T J kernel alone on one B200.

Per case one JSON line: time per launch (CUDA events on the library stream, L2 flushed between launches), queries/s,
algorithmic bytes (k-NN 16 N + 56 Q; J^T J = bytes of the factor records actually read + the partial sums)
and the fraction of the measured HBM peak.  --jtj-batch also times the J^T J kernel on batches of up to 4M factors
(many frames' correspondences in one launch), the size at which its traffic exceeds launch latency.

  python tools/sweep.py [--quick] [--out gpurun_out/sweep.jsonl]
"""
from __future__ import annotations

import argparse
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def peak_gbs():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    try:
        return float(json.load(open(p))["hbm_gbs"])
    except Exception:
        return 6650.0


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--quick", action="store_true")
    ap.add_argument("--out", default=None)
    ap.add_argument("--reps", type=int, default=10)
    ap.add_argument("--cell", type=float, default=0.0, help="voxel edge (0 = library default)")
    ap.add_argument("--jtj-batch", action="store_true")
    ap.add_argument("--check", action="store_true", help="verify a query sample against the CPU oracle (bit-exact)")
    args = ap.parse_args()

    import torch
    import ilsm_b200 as ilsm

    ilsm._build.build()
    dev = torch.device("cuda:0")
    ctx = ilsm.Context(0)
    ext = torch.cuda.ExternalStream(ctx.stream_ptr, device=dev)
    from bench import L2Flush
    flush = L2Flush(torch, dev)  # 256 MiB write + 256 MiB read: cold and clean L2 before every timed launch
    peak = peak_gbs()
    opts = ilsm.default_opts()
    out = open(args.out, "w") if args.out else None

    def emit(rec):
        s = json.dumps(rec)
        print(s, flush=True)
        if out:
            out.write(s + "\n")
            out.flush()

    def timed(fn, reps):
        ts = []
        with torch.cuda.stream(ext):
            for _ in range(3):
                fn()
            ctx.sync()
            for _ in range(reps):
                flush.zero_()
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record(ext)
                fn()
                e1.record(ext)
                ctx.sync()
                ts.append(e0.elapsed_time(e1))
        return float(np.median(ts)), float(np.min(ts))

    Ns = [100_000, 2_000_000] if args.quick else [100_000, 250_000, 500_000, 1_000_000, 2_000_000]
    Qs = [4096, 65536]
    S = ilsm.synth
    for N in Ns:
        c = S.config1(n_map=N)
        scene = c["scene"]
        m_all = np.concatenate([c["map_corner"], c["map_surf"]]).astype(np.float32)
        n_map = len(m_all)
        h_map = np.zeros((n_map, 4), np.float32)
        h_map[:, :3] = m_all[:, :3]
        d_map = torch.from_numpy(h_map).to(dev)
        gmap = ctx.new_map()
        t_build, _ = timed(lambda: gmap.build_dev(d_map.data_ptr(), n_map, 16, args.cell).join(), args.reps)
        emit({"case": "map_build", "N": n_map, "ms": t_build, "algorithmic_bytes": 48 * n_map,
              "GBs": 48 * n_map / t_build / 1e6, "frac": 48 * n_map / t_build / 1e6 / peak})
        # separate corner / surf maps for the fused associate kernel
        hc, hs = np.zeros((len(c["map_corner"]), 4), np.float32), np.zeros((len(c["map_surf"]), 4), np.float32)
        hc[:, :3], hs[:, :3] = c["map_corner"][:, :3], c["map_surf"][:, :3]
        d_mc, d_ms = torch.from_numpy(hc).to(dev), torch.from_numpy(hs).to(dev)
        mc, ms = ctx.new_map(), ctx.new_map()
        mc.build_dev(d_mc.data_ptr(), len(hc), 16, args.cell)
        ms.build_dev(d_ms.data_ptr(), len(hs), 16, args.cell)
        pose_true = np.concatenate([c["q_true"], c["t_true"]])
        d_pose = torch.from_numpy(pose_true).to(dev)
        d_out32 = torch.zeros(32, dtype=torch.float64, device=dev)
        for Q in Qs:
            # queries: Q = 65536 is the full organised frame (no-returns included, as the reference would feed them);
            # Q = 4096 a typical down-sampled stack: frame points sub-sampled
            cloud = c["cloud"][:, :3].astype(np.float32)
            rng = np.random.default_rng(Q)
            R = S.quat_to_mat(c["q_true"])
            world = (cloud.astype(np.float64) @ R.T + c["t_true"]).astype(np.float32)
            hit = np.linalg.norm(cloud, axis=1) > 0.1
            if Q >= len(cloud):
                qs = world[np.arange(Q) % len(cloud)]
            else:
                qs = world[hit][rng.choice(int(hit.sum()), Q, replace=int(hit.sum()) < Q)]
            hq = np.zeros((Q, 4), np.float32)
            hq[:, :3] = qs
            d_q = torch.from_numpy(hq).to(dev)
            d_idx = torch.empty((Q, 5), dtype=torch.int32, device=dev)
            d_d2 = torch.empty((Q, 5), dtype=torch.float32, device=dev)
            for md, tag in ((0.0, "exact"), (1.0, "gated(max_dist=1m)")):
                t_knn, t_min = timed(lambda: gmap.knn_dev(d_q.data_ptr(), Q, 16, 5, md, d_idx.data_ptr(), d_d2.data_ptr()),
                                     args.reps)
                byt = 16 * n_map + 56 * Q
                emit({"case": "knn5", "mode": tag, "N": n_map, "Q": Q, "ms": t_knn, "ms_min": t_min,
                      "queries_per_s": Q / t_knn * 1e3, "algorithmic_bytes": byt, "GBs": byt / t_knn / 1e6,
                      "frac": byt / t_knn / 1e6 / peak})
            if args.check:
                import oracle
                ctx.sync()
                sel = rng.choice(Q, 512, replace=False)
                gi, gd = d_idx.cpu().numpy()[sel], d_d2.cpu().numpy()[sel]
                ri, rd = oracle.knn_kdtree(h_map[:, :3].copy(), hq[sel, :3].copy(), 5)
                near = rd[:, 4] < 1.0  # gated run: exact within the gate
                assert np.array_equal(gi[near], ri[near]) and np.array_equal(gd[near], rd[near]), "k-NN parity"
                emit({"case": "knn5_check", "N": n_map, "Q": Q, "checked": int(near.sum()), "ok": True})
            # fused associate (pose transform + 5-NN + fit): sensor-frame points, half corner-tagged / half surf-tagged
            sens = np.zeros((Q, 4), np.float32)
            if Q >= len(cloud):
                sens[:, :3] = cloud[np.arange(Q) % len(cloud)]
            else:
                sens[:, :3] = cloud[hit][rng.choice(int(hit.sum()), Q, replace=int(hit.sum()) < Q)]
            ncq = Q // 8
            d_c, d_s = torch.from_numpy(sens[:ncq].copy()).to(dev), torch.from_numpy(sens[ncq:].copy()).to(dev)
            ctx.register_dev  # noqa: B018 (API presence)
            t_as, _ = timed(lambda: ctx.associate_dev(mc, ms, d_c.data_ptr(), ncq, d_s.data_ptr(), Q - ncq, 16,
                                                      d_pose.data_ptr(), opts), args.reps)
            byt = 16 * n_map + Q * (16 + 40 + 84)
            emit({"case": "associate", "N": n_map, "Q": Q, "ms": t_as, "points_per_s": Q / t_as * 1e3,
                  "algorithmic_bytes": byt, "GBs": byt / t_as / 1e6, "frac": byt / t_as / 1e6 / peak})
            t_j, t_jmin = timed(lambda: ctx.eval_normal_eq_dev(d_pose.data_ptr(), d_out32.data_ptr()), args.reps)
            byt = Q * 52 + ncq * 32 + ((Q + 383) // 384) * 256
            emit({"case": "jtj", "N": n_map, "Q": Q, "ms": t_j, "ms_min": t_jmin, "factors_per_s": Q / t_j * 1e3,
                  "algorithmic_bytes": byt, "GBs": byt / t_j / 1e6, "frac": byt / t_j / 1e6 / peak})
        if args.jtj_batch and N == Ns[-1]:
            for Qb in (262_144, 1_048_576, 4_194_304):
                sens = np.zeros((Qb, 4), np.float32)
                sens[:, :3] = cloud[np.arange(Qb) % len(cloud)]
                ncq = Qb // 8
                d_c, d_s = torch.from_numpy(sens[:ncq].copy()).to(dev), torch.from_numpy(sens[ncq:].copy()).to(dev)
                with torch.cuda.stream(ext):
                    ctx.associate_dev(mc, ms, d_c.data_ptr(), ncq, d_s.data_ptr(), Qb - ncq, 16, d_pose.data_ptr(), opts)
                    ctx.sync()
                t_j, t_jmin = timed(lambda: ctx.eval_normal_eq_dev(d_pose.data_ptr(), d_out32.data_ptr()), args.reps)
                byt = Qb * 52 + ncq * 32 + 148 * 256  # slots read once (b only for corner slots) + per-CTA partial sums
                emit({"case": "jtj_batched", "Q": Qb, "ms": t_j, "ms_min": t_jmin, "factors_per_s": Qb / t_j * 1e3,
                      "algorithmic_bytes": byt, "GBs": byt / t_j / 1e6, "frac": byt / t_j / 1e6 / peak})
        mc.close(), ms.close(), gmap.close()
    ctx.close()


if __name__ == "__main__":
    main()
