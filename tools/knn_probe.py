#!/usr/bin/env python
"""Config-3 k-NN timing probe: N-point map, the whole 65536-ray frame as queries, exact and gated, per-query vs binned path.
  python tools/knn_probe.py [N]"""
import json, os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import ilsm_b200 as ilsm

N = int(sys.argv[1]) if len(sys.argv) > 1 else 2_000_000
S = ilsm.synth
c = S.config1(n_map=N)
m = np.zeros((len(c["map_corner"]) + len(c["map_surf"]), 4), np.float32)
m[:, :3] = np.concatenate([c["map_corner"], c["map_surf"]])[:, :3]
R = S.quat_to_mat(c["q_true"])
w = np.zeros((65536, 4), np.float32)
w[:, :3] = (c["cloud"][:, :3].astype(np.float64) @ R.T + c["t_true"]).astype(np.float32)
dev = torch.device("cuda:0")
out = {"N": len(m), "Q": 65536}
for name, env, grp in (("per_query", "1000000000", "8"), ("binned_g4", "1", "4"), ("binned_g8", "1", "8"), ("binned_g16", "1", "16")):
    os.environ["ILSM_KNN_BINNED_MIN"] = env
    os.environ["ILSM_KNN_GROUP"] = grp
    ctx = ilsm.Context(0)
    ext = torch.cuda.ExternalStream(ctx.stream_ptr, device=dev)
    d_m, d_q = torch.from_numpy(m).to(dev), torch.from_numpy(w).to(dev)
    d_idx = torch.empty((65536, 5), dtype=torch.int32, device=dev)
    d_d2 = torch.empty((65536, 5), dtype=torch.float32, device=dev)
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
    gm = ctx.new_map()
    with torch.cuda.stream(ext):
        gm.build_dev(d_m.data_ptr(), len(m), 16)
        for md in (0.0, 1.0):
            ts = []
            for it in range(13):
                flush.zero_()
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record(ext)
                gm.knn_dev(d_q.data_ptr(), 65536, 16, 5, md, d_idx.data_ptr(), d_d2.data_ptr())
                e1.record(ext)
                ctx.sync()
                if it >= 3:
                    ts.append(e0.elapsed_time(e1))
            out[f"{name}_maxdist{md}"] = {"ms_median": float(np.median(ts)), "ms_min": float(np.min(ts))}
        out[f"{name}_idx_sum"] = int(d_idx.to(torch.int64).sum().item())
    gm.close(); ctx.close()
print(json.dumps(out))
