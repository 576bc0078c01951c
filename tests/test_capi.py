"""CPU: the C-ABI shared library loads and exports every symbol include/fastdm_b200.h declares
(no compute calls -- there is no GPU in the build container)."""
import ctypes
import os
import re

from conftest import ROOT


def declared_symbols():
    text = open(os.path.join(ROOT, "include", "fastdm_b200.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(fdm_[a-z0-9_]+)\s*\(", text)))


def test_header_declares_expected_entry_points():
    syms = declared_symbols()
    for name in ("fdm_quant_fp8", "fdm_quant_int8", "fdm_rms_norm", "fdm_rope", "fdm_gelu_and_mul",
                 "fdm_gemm_fp8", "fdm_gemm_int8", "fdm_attn_fwd", "fdm_last_error"):
        assert name in syms


def test_library_exports_every_declared_symbol(lib):
    for name in declared_symbols():
        assert hasattr(lib, name), f"libfastdm_b200.so does not export {name}"


def test_python_signatures_cover_header(lib):
    from fastdm_b200 import _lib

    assert sorted(_lib.SIGNATURES) == declared_symbols()


def test_no_torch_in_abi():
    text = open(os.path.join(ROOT, "include", "fastdm_b200.h")).read()
    assert "torch" not in re.sub(r"/\*.*?\*/", "", text, flags=re.S).lower()
    assert "at::" not in text and "c10::" not in text


def test_version_and_error_strings(lib):
    assert lib.fdm_version().decode().startswith("fastdm_b200")
    assert isinstance(lib.fdm_last_error(), bytes)


def test_library_has_no_libcuda_or_torch_dependency(lib):
    import subprocess

    from fastdm_b200 import _lib

    out = subprocess.run(["ldd", _lib.LIB_PATH], capture_output=True, text=True).stdout
    assert "libtorch" not in out and "libc10" not in out
