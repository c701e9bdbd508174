"""The C-ABI library: builds, loads, and exports exactly what include/gedepth.h declares.  No GPU."""
import ctypes
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def lib():
    from gedepth_b200 import build, kernels
    build.build()
    return kernels.load()


def _declared():
    src = open(os.path.join(ROOT, "include", "gedepth.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return re.findall(r"\b(?:int|const char\*)\s+(ged_\w+)\s*\(([^;]*?)\)\s*;", src, flags=re.S)


def test_header_symbols_are_exported(lib):
    decls = _declared()
    assert len(decls) >= 25
    for name, _ in decls:
        assert hasattr(lib, name), f"{name} declared in include/gedepth.h but not exported"


def test_binding_table_matches_header(lib):
    from gedepth_b200 import kernels
    decls = dict(_declared())
    for name, argtypes in kernels.SIGNATURES.items():
        assert name in decls, name
        nargs = len([a for a in decls[name].split(",") if a.strip() and a.strip() != "void"])
        assert nargs == len(argtypes), (name, nargs, len(argtypes))
    assert set(decls) - set(kernels.SIGNATURES) == {"ged_version", "ged_arch", "ged_msda_atomic_probe"}


def test_version_and_arch(lib):
    assert lib.ged_version() >= 100
    assert lib.ged_arch() == b"sm_100a"


def test_no_torch_types_in_abi():
    hdr = re.sub(r"/\*.*?\*/", "", open(os.path.join(ROOT, "include", "gedepth.h")).read(), flags=re.S)
    assert "torch" not in hdr.lower() and "at::" not in hdr and "Tensor" not in hdr


def test_null_arguments_are_rejected_without_a_gpu(lib):
    # argument validation happens before any CUDA call
    assert lib.ged_ge_vanilla_fwd(None, 0, None, None, None, 1, 4, 4, 2, 2, None) == -1
    assert lib.ged_gemm_tf32(None, 0, None, 0, None, 0, 1, 1, 4, None, 0, 0.0, None, None, 1, None, 0.0, 0, None, None) == -1


def test_sass_has_tcgen05_and_tma():
    so = os.path.join(ROOT, "gedepth_b200", "libgedepth_sm100.so")
    import subprocess
    out = subprocess.run(["cuobjdump", "-sass", so], capture_output=True, text=True).stdout
    assert "UTCHMMA" in out or "UTCQMMA" in out     # tcgen05.mma
    assert "UTMALDG" in out                           # cp.async.bulk.tensor
    assert "LDTM" in out                              # tcgen05.ld
