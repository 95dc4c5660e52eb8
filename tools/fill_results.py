import json, re, sys
def load(p): return json.loads(open(p).read().strip().splitlines()[-1])
b1, b2, b8 = load('profiles/r02_bench_n1.json'), load('profiles/r02_bench_n2.json'), load('profiles/r02_bench_n8.json')
def k(x, d=1): return f"{x/1e3:.{d}f} k"
c2, c3, c4, c5 = b1['config2'], b1['config3'], b1['config4'], b1['config5']
fe = b1['config1_with_frontend']
R = {
 'V1': k(b1['value'], 2), 'MS1': f"{1e3*b1['ms_per_step']:.1f}", 'E1': k(b1['e2e']['value'], 2), 'EMS1': f"{1e3*b1['e2e']['ms_per_step']:.0f}",
 'CPU1': f"{b1['cpu_baseline']['value']:.1f}", 'R1': f"{b1['e2e']['value']/b1['cpu_baseline']['value']:.0f}",
 'FE1': f"{1e3*fe['ms_per_step']:.0f}", 'FEE1': f"{1e3*fe['e2e']['ms_per_step']:.0f}",
 'S2': k(c2['value'], 2), 'SM2': f"{c2['ms_per_frame']:.3f}", 'P2': k(c2['pipelined_mapping_stage']['value'], 2),
 'T2': k(c2['three_stage_pipeline']['value'], 2), 'CPU2': f"{c2['cpu_baseline']['value']:.1f}",
 'R2': f"{c2['three_stage_pipeline']['value']/c2['cpu_baseline']['value']:.0f}",
 'D2': f"{c2['max_pose_diff_vs_oracle']['m']:.1e}", 'DR2': f"{c2['max_pose_diff_vs_oracle']['rad']:.1e}",
 'DA2': f"{c2['max_pose_diff_vs_oracle']['m_after_first_roll']:.1e}", 'NA2': str(c2['max_pose_diff_vs_oracle']['frames_after_first_roll']),
 'ROLL': str(c2['first_roll_frame']),
 'K3': f"{c3['knn5_exact']['ms']:.3f}", 'KG3': f"{c3['knn5_gated']['ms']:.3f}", 'KQ3': f"{c3['knn5_exact']['queries_per_s']/1e6:.0f}",
 'KF3': f"{100*c3['knn5_exact']['frac']:.1f} %",
 'J3': f"{1e3*c3['jtj']['ms']:.0f}", 'JG3': f"{c3['jtj']['achieved_GBs']/1e3:.2f}", 'JF3': f"{100*c3['jtj']['frac']:.0f} %",
 'CK3': f"{b1['cpu_baselines']['knn_all_cores']['value']/1e6:.0f}",
 'W4': k(c4['weak']['value'], 2), 'ST4': k(c4['strong']['value'], 1),
 'W4_2': k(b2['config4']['weak']['value'], 2), 'ST4_2': k(b2['config4']['strong']['value'], 1),
 'W4_8': k(b8['config4']['weak']['value'], 1), 'ST4_8': k(b8['config4']['strong']['value'], 1),
 'V2': k(b2['value'], 1), 'V8': k(b8['value'], 1), 'EFF8': f"{b8['value']/8/b1['value']:.2f}",
 'C5': k(c5['value'], 1), 'C5_2': k(b2['config5']['value'], 1), 'C5_8': k(b8['config5']['value'], 1),
 'C5E2': f"{b2['config5']['value']/2/c5['value']:.2f}", 'C5E8': f"{b8['config5']['value']/8/c5['value']:.2f}",
 'C5R': k(c5['reference_exact_candidates']['value'], 1),
}
for p in ('DESIGN.md', 'README.md'):
    s = open(p).read()
    for key, v in R.items(): s = s.replace('@' + key + '@', v)
    left = re.findall(r'@[A-Z0-9_]+@', s)
    assert not left, (p, left)
    open(p, 'w').write(s)
print(json.dumps(R, indent=0))
