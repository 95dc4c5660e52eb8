import sys, numpy as np
sys.path.insert(0, '.')
import torch, ilsm_b200 as ilsm
S = ilsm.synth
ctx = ilsm.Context(0)
sc = ilsm.ScanContextDb(ctx)
N = 100_000
for a in range(0, N, 20_000):
    sc.add(S.sc_database_range(a, a + 20_000, N))
q, _, _ = S.sc_chunked_queries(N, 8)
approx, sh = sc.prefilter_debug(q.reshape(8, 1200), N - 50)
approx = np.asarray(approx).reshape(8, -1)
print("hard-flagged per query", (approx == -1).sum(axis=1), "dual", (approx <= -2).sum(axis=1), "of", approx.shape[1])
for qi in range(8):
    a = approx[qi]; a = a[a >= 0]
    T = np.partition(a, 9)[9]
    print(qi, "T", T, "within T+2eps", int((a <= T + 3e-3).sum()))
