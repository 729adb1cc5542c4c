"""C-ABI surface: the library loads and exports every symbol include/satmc.h declares (no GPU needed)."""
import ctypes
import os
import re

import pytest


def test_library_exports_every_declared_symbol(satmc):
    lib = satmc.load_library()
    header = open(os.path.join(satmc.INCLUDE_DIR, "satmc.h")).read()
    declared = set(re.findall(r"\b(satmc_[a-z_0-9]+)\s*\(", header))
    declared -= {"satmc_pair", "satmc_ctx", "satmc_status"}
    assert declared, "no declarations parsed"
    for name in sorted(declared):
        assert hasattr(lib, name), f"libsatmc.so does not export {name}"
    assert declared == set(satmc.ABI_SYMBOLS), declared ^ set(satmc.ABI_SYMBOLS)


def test_pair_layout_is_48_bytes(satmc):
    assert satmc.PAIR_DTYPE.itemsize == 48
    assert satmc.PAIR_DTYPE.names[:3] == ("rx", "ry", "rtheta")


def test_version_string(satmc):
    lib = satmc.load_library()
    assert b"sm_100a" in lib.satmc_version()


def test_no_cpu_fallback_without_gpu(satmc):
    """Without a CUDA device the context must fail loudly (never compute on the CPU)."""
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    with pytest.raises(satmc.SatmcError) as e:
        satmc.Context(0)
    assert e.value.code == -2          # SATMC_ERR_NO_DEVICE


def test_product_does_not_reference_oracle():
    """The product path must not import, link or call anything under oracle/."""
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    pkg = os.path.join(root, "convex-2d-gpu-collision-detection_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".cpp", ".h", ".hpp")) or f == "Makefile":
                text = open(os.path.join(dirpath, f), errors="ignore").read()
                assert "sat_oracle" not in text and "oracle/" not in text and "orc_" not in text, \
                    f"{f} references the oracle"


def test_header_is_valid_c_and_links_from_c(satmc, tmp_path):
    """include/satmc.h compiles as C99 and a plain C client links against libsatmc.so (no torch, no C++)."""
    import subprocess
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    libdir = os.path.dirname(satmc.LIB_PATH)
    exe = str(tmp_path / "abi_example")
    subprocess.check_call(["gcc", "-std=c99", "-Wall", "-Wextra", "-Werror", "-pedantic", "-I", satmc.INCLUDE_DIR,
                           os.path.join(root, "tests", "c", "abi_example.c"), "-o", exe, "-L", libdir, "-lsatmc", "-lm",
                           f"-Wl,-rpath,{libdir}"])
    exe2 = str(tmp_path / "group_example")
    subprocess.check_call(["gcc", "-std=c99", "-Wall", "-Wextra", "-Werror", "-pedantic", "-I", satmc.INCLUDE_DIR,
                           os.path.join(root, "tests", "c", "group_example.c"), "-o", exe2, "-L", libdir, "-lsatmc", "-lm",
                           f"-Wl,-rpath,{libdir}"])
    import torch
    for e in (exe, exe2):
        r = subprocess.run([e], capture_output=True, text=True)
        if torch.cuda.is_available():
            assert r.returncode == 0, r.stdout + r.stderr
        else:
            assert r.returncode == 77, r.stdout + r.stderr   # fails loudly without a GPU
