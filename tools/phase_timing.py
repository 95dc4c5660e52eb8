"""Profiling aid: build a copy of the library with -DILSM_DEBUG_TIMING and print clock64() phase timings of the
solve and associate kernels on config 1 (run on the GPU box)."""
import ctypes as C, os, subprocess, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import ilsm_b200 as ilsm
from ilsm_b200 import _build
dbg = os.path.join(_build.HERE, "libilsm_cuda_dbg.so")
_build.build(force=True, extra_flags=["-DILSM_DEBUG_TIMING"], lib=dbg, obj_dir=_build.OBJ + "_dbg")
ilsm.binding._lib = None
lib = ilsm.load_library(dbg)
lib.ilsm_debug_stamps.argtypes = [C.c_void_p, C.c_void_p]
c = ilsm.synth.config1(100_000)
ctx = ilsm.Context(0)
mc = ctx.new_map().set_input_cloud(c["map_corner"])
ms = ctx.new_map().set_input_cloud(c["map_surf"])
for rep in range(3):
    q, t, r = ctx.register(mc, ms, c["corner"], c["surf"], c["q0"], c["t0"])
    buf = np.zeros(64, np.int64)
    lib.ilsm_debug_stamps(ctx._h, buf.ctypes.data_as(C.c_void_p))
    print("---- rep", rep, "iterations", r.pass_[1].iterations)
    s = buf[:48].reshape(6, 8)
    names = ["eval", "warp-reduce+sync", "cluster-sync", "dsmem-sum+sync", "LM", "sync"]
    for e in range(6):
        if s[e, 0] == 0:
            continue
        d = np.diff(s[e, :7])
        print(f"eval {e}: " + "  ".join(f"{n}={int(v)}" for n, v in zip(names, d)) + f"  total={int(s[e,6]-s[e,0])}")
    for k, nm in ((48, "corner warp"), (56, "surf warp")):
        a = buf[k:k + 4]
        print(nm, "pose+transform", int(a[1] - a[0]), "knn", int(a[2] - a[1]), "fit", int(a[3] - a[2]))
