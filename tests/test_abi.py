"""CPU-side checks of the drop-in boundary: the C-ABI library builds, loads and
exports every symbol include/superdiff_b200.h declares (no compute calls)."""
import ctypes
import os
import re

from super_diffusion_b200 import _lib

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared_symbols():
    src = open(os.path.join(ROOT, "include", "superdiff_b200.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(sd_[a-z0-9_]+)\s*\(", src)))


def test_header_symbols_are_bound_and_exported(lib):
    declared = _declared_symbols()
    assert declared, "no symbols parsed from the header"
    assert sorted(_lib.SIGNATURES) == declared
    raw = ctypes.CDLL(_lib.LIB_PATH)
    for name in declared:
        assert hasattr(raw, name), f"{name} declared in the header but not exported by the .so"


def test_version_and_error_string(lib):
    assert lib.sd_version() >= 100
    assert isinstance(lib.sd_last_error(), bytes)


def test_argument_validation_without_gpu(lib):
    # invalid arguments are rejected before any CUDA call, so this runs on CPU
    null = ctypes.c_void_p(0)
    sp = (ctypes.c_void_p * 1)(0)
    rc = lib.sd_step_vpsde(null, null, sp, 9, 1, 4, 0.0, 1.0, 1.0, 1e-3, null, null, 0, 0, 1.0, null, 0.0,
                           null, null, null, null)
    assert rc == -1 and b"M must be" in lib.sd_last_error()
    rc = lib.sd_step_edm_cfg(null, null, null, null, null, 1, 16, 1.0, -0.1, 7.5, 0.0, 1, 1.0, 0.0, 0.5,
                             null, null, null, null)
    assert rc == -1


def test_no_oracle_import_in_product():
    pkg = os.path.join(ROOT, "super_diffusion_b200")
    for dp, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                txt = open(os.path.join(dp, f)).read()
                assert not re.search(r"^\s*(from|import)\s+oracle\b", txt, flags=re.M), f


def test_examples_compile_as_plain_c(tmp_path):
    """include/superdiff_b200.h is a C header (extern "C" guards, no C++ types): the two example hosts build with gcc and
    link against the shared library.  Compile + link only; running them needs a GPU (tests/test_scorenet_native.py)."""
    import shutil
    import subprocess
    import pytest
    if shutil.which("gcc") is None or not os.path.exists("/usr/local/cuda/include/cuda_runtime_api.h"):
        pytest.skip("needs gcc and the CUDA runtime headers")
    for name in ("native_forward", "native_sampler"):
        r = subprocess.run(["gcc", "-O1", "-Wall", "-Werror", "-I", os.path.join(ROOT, "include"), "-I", "/usr/local/cuda/include",
                            os.path.join(ROOT, "examples", name + ".c"), "-o", str(tmp_path / name),
                            "-L", os.path.join(ROOT, "super_diffusion_b200"), "-lsuperdiff_b200", "-L", "/usr/local/cuda/lib64",
                            "-lcudart"], capture_output=True, text=True)
        assert r.returncode == 0, r.stderr
