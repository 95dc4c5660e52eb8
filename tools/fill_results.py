#!/usr/bin/env python
"""Regenerates the measured-results tables of DESIGN.md (section 5) and README.md from the committed bench lines
(profiles/r02_bench_n{1,2,4,8}.json):  python tools/fill_results.py"""
import json
import os
import re

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def load(n):
    return json.loads(open(os.path.join(ROOT, "profiles", f"r02_bench_n{n}.json")).read().strip().splitlines()[-1])


def k(x, d=1):
    return f"{x / 1e3:.{d}f} k"


b1, b2, b4, b8 = load(1), load(2), load(4), load(8)
c2, c3, c4, c5, fe = b1["config2"], b1["config3"], b1["config4"], b1["config5"], b1["config1_with_frontend"]
pd = c2["max_pose_diff_vs_oracle"]
R = {
    "V1": k(b1["value"], 2), "MS1": f"{1e3 * b1['ms_per_step']:.1f}", "E1": k(b1["e2e"]["value"], 2), "EMS1": f"{1e3 * b1['e2e']['ms_per_step']:.0f}",
    "CPU1": f"{b1['cpu_baseline']['value']:.1f}", "R1": f"{b1['e2e']['value'] / b1['cpu_baseline']['value']:.0f}",
    "FE1": f"{1e3 * fe['ms_per_step']:.0f}", "FEE1": f"{1e3 * fe['e2e']['ms_per_step']:.0f}",
    "S2": k(c2["value"], 2), "SM2": f"{c2['ms_per_frame']:.3f}", "L2": f"{c2['gpu_launches_per_frame']:.0f}", "P2": k(c2["pipelined_mapping_stage"]["value"], 2),
    "T2": k(c2["three_stage_pipeline"]["value"], 2), "TM2": f"{c2['three_stage_pipeline']['ms_per_call_median']:.2f}",
    "CPU2": f"{c2['cpu_baseline']['value']:.1f}", "R2": f"{c2['three_stage_pipeline']['value'] / c2['cpu_baseline']['value']:.0f}",
    "D2": f"{pd['m']:.1e}", "DR2": f"{pd['rad']:.1e}", "DA2": f"{pd['m_after_first_roll']:.1e}", "NA2": str(pd["frames_after_first_roll"]),
    "ROLL": str(c2["first_roll_frame"]),
    "K3": f"{c3['knn5_exact']['ms']:.3f}", "KG3": f"{c3['knn5_gated']['ms']:.3f}", "KQ3": f"{c3['knn5_exact']['queries_per_s'] / 1e6:.0f}",
    "KF3": f"{100 * c3['knn5_exact']['frac']:.1f} %", "J3": f"{1e3 * c3['jtj']['ms']:.0f}", "JG3": f"{c3['jtj']['achieved_GBs'] / 1e3:.2f}",
    "JF3": f"{100 * c3['jtj']['frac']:.0f} %", "CK3": f"{b1['cpu_baselines']['knn_all_cores']['value'] / 1e6:.0f}",
    "W4": k(c4["weak"]["value"], 2), "ST4": k(c4["strong"]["value"], 1),
    "W4_2": k(b2["config4"]["weak"]["value"], 2), "ST4_2": k(b2["config4"]["strong"]["value"], 1),
    "W4_4": k(b4["config4"]["weak"]["value"], 1), "ST4_4": k(b4["config4"]["strong"]["value"], 1),
    "W4_8": k(b8["config4"]["weak"]["value"], 1), "ST4_8": k(b8["config4"]["strong"]["value"], 1),
    "V2": k(b2["value"], 1), "V4": k(b4["value"], 1), "V8": k(b8["value"], 1), "EFF8": f"{b8['value'] / 8 / b1['value']:.2f}",
    "E8": k(b8["e2e"]["value"], 1),
    "C5": k(c5["value"], 1), "C5_2": k(b2["config5"]["value"], 1), "C5_4": k(b4["config5"]["value"], 1), "C5_8": k(b8["config5"]["value"], 1),
    "C5E2": f"{b2['config5']['value'] / 2 / c5['value']:.2f}", "C5E4": f"{b4['config5']['value'] / 4 / c5['value']:.2f}",
    "C5E8": f"{b8['config5']['value'] / 8 / c5['value']:.2f}", "C5R": k(c5["reference_exact_candidates"]["value"], 1),
}

DESIGN_TABLE = """| Config | GPU | CPU (same run) | Notes |
|---|---|---|---|
| 1: frame vs 100 k map | **@V1@ reg/s** device-resident (@MS1@ µs), **@E1@ reg/s e2e** (@EMS1@ µs incl. 1.6 MB H2D) | @CPU1@ reg/s (1 thread, reference nanoflann tree) | ×@R1@ e2e.  Step = solve 2 × 18–19 µs, associate 2 × 7 µs, both builds 18 µs; all latency-bound (fp64 dependency chains at ~35 cycles per operation, §4 K3): `roofline.frac` 0.2 % (solve), 2–4 % (associate, build).  With its front end: @FE1@ µs device-timed, @FEE1@ µs e2e (CPU front end alone: 2.8–3.1 ms) |
| 2: full loop, 2000 frames | synchronous **@S2@ frames/s** (@SM2@ ms/frame, @L2@ launches), mapping as its own stage **@P2@**, three stages **@T2@** (@TM2@ ms per call median) | @CPU2@ frames/s (chained oracle, all 2000 frames) | ×@R2@.  Largest pose difference against the oracle over 2000 frames **@D2@ m** / @DR2@ rad, @DA2@ m over the @NA2@ frames after the in-loop window roll at frame @ROLL@; no capacity flags |
| 3: N = 2 M, Q = 65 536 | 5-NN exact **@K3@ ms**, gated @KG3@ ms (@KQ3@ M queries/s; 382 warp instructions per query); JᵀJ 4 M factors **@J3@ µs = @JG3@ TB/s = @JF3@ of the measured peak** | 16-core k-d tree: @CK3@ M queries/s | k-NN: @KF3@ of peak against `16·N + 56·Q`, but ncu sees only 4.8 MB of DRAM traffic — the search touches 5 of 32 MB of the map and runs out of L2; it is bound by instructions (§4 K1), and an exact search that skips most of the map cannot reach an HBM roofline defined on the whole map |
| 4: independent sequences | one sequence per GPU: @W4@ / @W4_2@ / @W4_4@ / **@W4_8@ frames/s** at 1 / 2 / 4 / 8 GPUs; 8 sequences over the GPUs: **@ST4@** / @ST4_2@ / @ST4_4@ / @ST4_8@ | — | no collective; one GPU already serves 8 sequences at 57 % of the 8-GPU rate (the per-frame chain is latency-bound, concurrent sequences fill the SMs).  Headline replicas @V1@ / @V2@ / @V4@ / **@V8@ reg/s** (efficiency @EFF8@), e2e @E8@ on 8 GPUs |
| 5: 100 k keyframes, top-10 | **@C5@ / @C5_2@ / @C5_4@ / @C5_8@ queries/s** at 1 / 2 / 4 / 8 GPUs (strong scaling @C5E2@ / @C5E4@ / @C5E8@), top-1 id + shift recovered for all queries; the reference's 10-candidate search: @C5R@ queries/s on one GPU | ≈ 1 query/s brute force | round 1: 2.16 k on one GPU, 7.4 k on 8 |
"""

README_TABLE = """| | GPU | CPU (same box, same run) |
|---|---|---|
| scan-to-map registration, OS0-64 frame vs 100 k-point map | @V1@ reg/s device-resident (@MS1@ µs), **@E1@ reg/s end to end** from host buffers | @CPU1@ reg/s (oracle on the reference's nanoflann tree, 1 thread) |
| full loop (front end + odometry + mapping), 2000-frame corridor | @S2@ frames/s synchronous; @P2@ with the mapping stage; **@T2@ as three stages**; @ST4@ with 8 concurrent sequences | @CPU2@ frames/s |
| JᵀJ evaluation, 4 M factors | @J3@ µs = @JG3@ TB/s = @JF3@ of the measured HBM peak (bulk-async-copy pipeline) | — |
| exact 5-NN, 2 M-point map, 65 536 queries | @K3@ ms exact / @KG3@ ms gated (queries binned by voxel; instruction-bound, runs out of L2) | @CK3@ M queries/s on 16 cores |
| ScanContext, 100 k keyframes, top-10 | **@C5@ queries/s** (f16 tensor-core prefilter + exact fp64 rescoring, results identical to an exact scan); @C5_8@ on 8 GPUs (NCCL all-gather behind the C ABI) | ≈ 1 query/s |
| 8 GPUs | @V8@ reg/s (replicas); @W4_8@ frames/s over 8 sequences | — |
"""


def fill(t):
    for key, v in R.items():
        t = t.replace("@" + key + "@", v)
    assert not re.findall(r"@[A-Z0-9_]+@", t), re.findall(r"@[A-Z0-9_]+@", t)
    return t


def replace_table(path, first_line, table):
    s = open(path).read()
    a = s.index(first_line)
    b = a
    while s[b:b + 1] == "|":  # the table ends at the first line that does not start with a bar
        b = s.index("\n", b) + 1
    open(path, "w").write(s[:a] + fill(table) + s[b:])


replace_table(os.path.join(ROOT, "DESIGN.md"), "| Config | GPU | CPU (same run) | Notes |", DESIGN_TABLE)
replace_table(os.path.join(ROOT, "README.md"), "| | GPU | CPU (same box, same run) |", README_TABLE)
print(json.dumps(R))
