"""Profiling aid: distribution of the e2e step (host wall clock) and of its parts, and of a bare pinned H2D copy of the
same bytes (run on the GPU box):  python tools/e2e_probe.py"""
import os, sys, time
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import ilsm_b200 as ilsm
import bench
bench.pin_affinity(torch, 0, 1)
c = ilsm.synth.config1(100_000)
def pad4(a):
    o = np.zeros((len(a), 4), np.float32); o[:, :3] = a[:, :3]; return o
pin = lambda a: torch.from_numpy(pad4(a)).pin_memory()
tmc, tms, tc, ts = pin(c["map_corner"]), pin(c["map_surf"]), pin(c["corner"]), pin(c["surf"])
mcn, msn, cn, sn = tmc.numpy(), tms.numpy(), tc.numpy(), ts.numpy()
ctx = ilsm.Context(0); mc, ms = ctx.new_map(), ctx.new_map()
ext = torch.cuda.ExternalStream(ctx.stream_ptr, device=torch.device("cuda:0"))
flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
ctx.set_async(True)
def pct(v):
    v = 1e6 * np.asarray(v)
    return "p10 %7.1f  p50 %7.1f  p90 %7.1f  max %7.1f us" % (np.percentile(v, 10), np.median(v), np.percentile(v, 90), v.max())
for do_flush in (True, False):
    T = {k: [] for k in ("build_c", "build_s", "register", "total")}
    with torch.cuda.stream(ext):
        for it in range(420):
            if do_flush:
                flush.zero_()
            ctx.sync(); t0 = time.perf_counter()
            mc.set_input_cloud(mcn); t1 = time.perf_counter()
            ms.set_input_cloud(msn); t2 = time.perf_counter()
            q, t, rep = ctx.register(mc, ms, cn, sn, c["q0"], c["t0"]); t3 = time.perf_counter()
            if it >= 20:
                T["build_c"].append(t1 - t0); T["build_s"].append(t2 - t1); T["register"].append(t3 - t2); T["total"].append(t3 - t0)
    print("flush between steps:", do_flush)
    for k, v in T.items():
        print(f"  {k:10s} {pct(v)}")
    tot = 1e6 * np.asarray(T["total"])
    print("  histogram of total (20 us bins from 100):", np.histogram(tot, bins=[0] + list(range(100, 320, 20)) + [1e9])[0].tolist())
# bare H2D of the same buffers on one stream
d = [torch.empty_like(x, device="cuda") for x in (tmc, tms, tc, ts)]
for do_flush in (True, False):
    tt = []
    for it in range(220):
        if do_flush:
            flush.zero_()
        torch.cuda.synchronize(); t0 = time.perf_counter()
        for a, b in zip(d, (tmc, tms, tc, ts)):
            a.copy_(b, non_blocking=True)
        torch.cuda.synchronize(); tt.append(time.perf_counter() - t0)
    print("bare H2D of %.2f MB, flush %s: %s" % (sum(x.numel() * 4 for x in (tmc, tms, tc, ts)) / 1e6, do_flush, pct(tt[20:])))
print("affinity", len(os.sched_getaffinity(0)), "cores")
