"""Profiling aid: host-side wall-clock breakdown of the e2e step (run on the GPU box)."""
import os, sys, time
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import ilsm_b200 as ilsm
c = ilsm.synth.config1(100_000)
def pad4(a):
    o = np.zeros((len(a), 4), np.float32); o[:, :3] = a[:, :3]; return o
pin = lambda a: torch.from_numpy(pad4(a)).pin_memory().numpy()
mcn, msn, cn, sn = pin(c["map_corner"]), pin(c["map_surf"]), pin(c["corner"]), pin(c["surf"])
ctx = ilsm.Context(0); mc, ms = ctx.new_map(), ctx.new_map()
ctx.set_async(True)
T = {k: [] for k in ("build_c", "build_s", "register", "total")}
for it in range(60):
    ctx.sync(); t0 = time.perf_counter()
    mc.set_input_cloud(mcn); t1 = time.perf_counter()
    ms.set_input_cloud(msn); t2 = time.perf_counter()
    q, t, rep = ctx.register(mc, ms, cn, sn, c["q0"], c["t0"]); t3 = time.perf_counter()
    if it >= 10:
        T["build_c"].append(t1 - t0); T["build_s"].append(t2 - t1); T["register"].append(t3 - t2); T["total"].append(t3 - t0)
for k, v in T.items():
    print(f"{k:10s} median {1e6*np.median(v):8.1f} us  min {1e6*np.min(v):8.1f}")
# raw H2D bandwidth from pinned memory
d = torch.empty(len(msn) * 4, dtype=torch.float32, device="cuda")
src = torch.from_numpy(msn).pin_memory()
for n in range(3):
    torch.cuda.synchronize(); t0 = time.perf_counter(); d.copy_(src.view(-1), non_blocking=True); torch.cuda.synchronize(); dt = time.perf_counter() - t0
print(f"H2D {src.numel()*4/1e6:.2f} MB in {dt*1e6:.1f} us = {src.numel()*4/dt/1e9:.1f} GB/s")
# blocking mode for comparison
ctx.set_async(False)
tt = []
for it in range(30):
    ctx.sync(); t0 = time.perf_counter(); mc.set_input_cloud(mcn); ms.set_input_cloud(msn); ctx.register(mc, ms, cn, sn, c["q0"], c["t0"]); tt.append(time.perf_counter() - t0)
print(f"blocking total median {1e6*np.median(tt[5:]):.1f} us")
