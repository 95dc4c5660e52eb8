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
