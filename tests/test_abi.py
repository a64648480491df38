"""CPU-side checks of the drop-in boundary: the library loads, exports exactly what
include/pantax_gpu.h declares, and refuses to run without a GPU (no fallback)."""
import ctypes as C
import os
import re

import pytest

import pantax_b200
from pantax_b200 import _lib, build

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HEADER = os.path.join(ROOT, "include", "pantax_gpu.h")


def header_functions():
    src = open(HEADER).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(ptx_[a-z0-9_]+)\s*\(", src)))


def test_library_builds_and_loads():
    path = build.build()
    assert os.path.exists(path)
    L = pantax_b200.load_library()
    assert b"sm_100a" in L.ptx_version()


def test_every_declared_symbol_is_exported_and_bound():
    L = C.CDLL(build.build())
    declared = header_functions()
    assert len(declared) >= 35
    for name in declared:
        assert hasattr(L, name), f"{name} declared in pantax_gpu.h but not exported"
    assert sorted(_lib.SIGNATURES) == declared, "ctypes binding and header diverge"


def test_header_cites_the_reference_interface_it_replaces():
    src = open(HEADER).read()
    for cite in ("rcls.rs:119-146", "profile.rs:787-919", "profile.rs:658-740", "profile.rs:2705-2729", "gaf_filter.rs:44-97",
                 "zip.rs:236-262", "profile.rs:361-463"):
        assert cite in src


def test_no_cpu_fallback_without_device():
    import torch

    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    from pantax_b200 import api

    with pytest.raises(pantax_b200.PantaxGpuError) as e:
        api.PantaxGpu(0)
    assert e.value.name == "PTX_E_CUDA"


def test_product_never_imports_the_oracle():
    pkg = os.path.join(ROOT, "pantax_b200")
    for dirpath, _d, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h", ".cpp")):
                txt = open(os.path.join(dirpath, f), errors="ignore").read()
                assert "oracle" not in txt.replace("oracles are the checkers", ""), f"{f} mentions the oracle"
