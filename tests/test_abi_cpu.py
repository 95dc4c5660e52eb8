"""The C-ABI library builds for sm_100a, loads without a GPU and exports every symbol include/ilsm.h declares."""
import ctypes
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared_symbols():
    txt = open(os.path.join(ROOT, "include", "ilsm.h")).read()
    return sorted(set(re.findall(r"ILSM_API[^;(]*?\b(ilsm_\w+)\s*\(", txt)))


def test_header_declares_entry_points():
    syms = declared_symbols()
    for s in ("ilsm_create", "ilsm_map_build", "ilsm_knn", "ilsm_register", "ilsm_eval_normal_eq", "ilsm_associate"):
        assert s in syms


def test_library_exports_every_declared_symbol(ilsm):
    lib = ctypes.CDLL(ilsm._build.LIB)
    for s in declared_symbols():
        assert hasattr(lib, s), f"{s} declared in include/ilsm.h but not exported"
    assert lib.ilsm_abi_version() == 1


def test_header_compiles_as_plain_c(tmp_path):
    src = tmp_path / "t.c"
    src.write_text('#include "ilsm.h"\nint main(void){ilsm_reg_opts o; (void)o; return sizeof(ilsm_factor)==80?0:1;}\n')
    import subprocess
    exe = tmp_path / "t"
    subprocess.run(["gcc", "-std=c99", "-Wall", "-Werror", "-I", os.path.join(ROOT, "include"), str(src), "-o", str(exe)],
                   check=True)
    assert subprocess.run([str(exe)]).returncode == 0


def test_struct_layouts_match_ctypes(ilsm):
    from ilsm_b200 import binding as b
    assert ctypes.sizeof(b.RegOpts) == 48
    assert ctypes.sizeof(b.SolveSummary) == 48
    assert ctypes.sizeof(b.RegReport) == 8 + 8 * 48
    assert b.FACTOR_DTYPE.itemsize == 80


def test_every_struct_layout_matches_the_c_header(ilsm, tmp_path):
    """sizeof of every struct of include/ilsm.h as gcc lays it out == the ctypes / numpy mirror in binding.py."""
    import subprocess
    from ilsm_b200 import binding as b
    pairs = {"ilsm_reg_opts": b.RegOpts, "ilsm_solve_summary": b.SolveSummary, "ilsm_reg_report": b.RegReport,
             "ilsm_feature_counts": b.FeatureCounts, "ilsm_features": b.Features, "ilsm_cubemap_stats": b.CubeMapStats,
             "ilsm_slam_stats": b.SlamStats, "ilsm_ground_opts": b.GroundOpts, "ilsm_ground_info": b.GroundInfo,
             "ilsm_mapopt_stats": b.MapOptStats, "ilsm_pc2_layout": b.Pc2Layout}
    names = list(pairs) + ["ilsm_dmatch", "ilsm_factor"]
    src = tmp_path / "s.c"
    src.write_text('#include <stdio.h>\n#include "ilsm.h"\nint main(void){' +
                   "".join(f'printf("%zu\\n", sizeof({n}));' for n in names) + "return 0;}\n")
    exe = tmp_path / "s"
    subprocess.run(["gcc", "-std=c99", "-I", os.path.join(ROOT, "include"), str(src), "-o", str(exe)], check=True)
    sizes = [int(x) for x in subprocess.run([str(exe)], capture_output=True, text=True).stdout.split()]
    got = dict(zip(names, sizes))
    for n, t in pairs.items():
        assert got[n] == ctypes.sizeof(t), (n, got[n], ctypes.sizeof(t))
    assert got["ilsm_dmatch"] == b.DMATCH_DTYPE.itemsize == 16 and got["ilsm_factor"] == b.FACTOR_DTYPE.itemsize


def test_null_handles_are_rejected_not_dereferenced(ilsm):
    """Argument validation runs before any CUDA call: usable without a GPU, negative status + error text."""
    lib = ilsm.load_library()
    n = ctypes.c_int(0)
    assert lib.ilsm_map_build(None, None, 0, 16, 0.0) < 0
    assert lib.ilsm_knn(None, None, 0, 16, 5, 0.0, None, None) < 0
    assert lib.ilsm_register(None, None, None, None, 0, None, 0, 16, None, None, None, None) < 0
    assert lib.ilsm_voxelgrid(None, None, 0, 16, 0.4, None, ctypes.byref(n)) < 0
    assert lib.ilsm_slam_frame(None, None, 0, 16, 1, None, None, None, None, None) < 0
    assert lib.ilsm_mapopt_frame(None, None, 0, 16, None, 0, 16, None, None, None, None, None, None) < 0
    assert lib.ilsm_orb_match(None, None, 0, None, 0, 32, 1, 0.3, None, ctypes.byref(n), None, ctypes.byref(n)) < 0
    assert lib.ilsm_ground_extract(None, None, 0, 16, None, None, 0, ctypes.byref(n), None, None) < 0
    assert b"null" in lib.ilsm_last_error() or b"bad" in lib.ilsm_last_error()


def test_no_cpu_fallback(ilsm):
    """Without a CUDA device the product path must fail loudly, never compute on the CPU."""
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    with pytest.raises(ilsm.IlsmError) as e:
        ilsm.Context(0)
    assert e.value.code == -3


def test_product_does_not_import_oracle():
    pkg = os.path.join(ROOT, "intensity_based_lidar_slam_for_me-_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".hpp", ".cpp", ".h")):
                txt = open(os.path.join(dirpath, f), errors="replace").read()
                assert "import oracle" not in txt and "from oracle" not in txt and "ilsm_oracle" not in txt, f
