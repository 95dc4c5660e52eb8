"""In-tree build of libilsm_cuda.so (sm_100a only).  nvcc cross-compiles without a GPU; the resulting .so is
git-ignored but travels to the GPU box with the repository snapshot."""
from __future__ import annotations

import os
import shutil
import subprocess

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
LIB = os.path.join(HERE, "libilsm_cuda.so")

NVCC_FLAGS = [
    "-O3", "-std=c++17", "-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo",
    "-Xcompiler", "-fPIC,-fvisibility=hidden",
]
# Parity with the reference's x86-64 SSE2 build (CMakeLists.txt:5-6, no FMA): files whose float results are compared
# bit for bit are compiled without a*b+c contraction.  registration.cu keeps its bit-exact parts (pose transform,
# float distances) in explicit round-to-nearest intrinsics and is compared with a tolerance elsewhere (fits, residuals,
# LM), so its fp64 arithmetic may use DFMA.
FMAD = {"registration.cu": "true"}
OBJ = os.path.join(HERE, "build")


def sources():
    return sorted(os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith(".cu"))


def _stale():
    if not os.path.exists(LIB):
        return True
    t = os.path.getmtime(LIB)
    deps = [os.path.join(CSRC, f) for f in os.listdir(CSRC)] + [os.path.join(HERE, "..", "include", "ilsm.h"), __file__]
    return any(os.path.getmtime(d) > t for d in deps)


def nvcc_path():
    return shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"


def build(force: bool = False, verbose: bool = False, extra_flags=(), lib: str = LIB, obj_dir: str = OBJ) -> str:
    """One nvcc -c per source (in parallel), then one link into libilsm_cuda.so (or `lib`, for instrumented variants)."""
    if not force and lib == LIB and not _stale():
        return lib
    os.makedirs(obj_dir, exist_ok=True)
    procs = []
    for src in sources():
        name = os.path.basename(src)
        obj = os.path.join(obj_dir, name[:-3] + ".o")
        cmd = ([nvcc_path()] + NVCC_FLAGS + list(extra_flags) + ["-fmad=" + FMAD.get(name, "false")] +
               (["-Xptxas", "-v"] if verbose else []) + ["-c", src, "-o", obj])
        procs.append((name, obj, subprocess.Popen(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)))
    objs = []
    for name, obj, p in procs:
        out, _ = p.communicate()
        if p.returncode != 0:
            raise RuntimeError(f"nvcc failed on {name}:\n{out}")
        if verbose:
            print(out)
        objs.append(obj)
    r = subprocess.run([nvcc_path(), "-shared", "-gencode", "arch=compute_100a,code=sm_100a", "-o", lib] + objs,
                       capture_output=True, text=True)
    if r.returncode != 0:
        raise RuntimeError("link failed:\n" + r.stdout + r.stderr)
    return lib
