"""The C-ABI library loads on a CPU-only box and exports every symbol that
include/ipxgpu.h declares (no compute calls here)."""

import ctypes
import os
import re

import pytest

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared_functions():
    text = open(os.path.join(REPO, "include", "ipxgpu.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    names = re.findall(r"\b(ipxgpu_[a-z0-9_]+)\s*\(", text)
    return sorted(set(n for n in names if n != "ipxgpu_interrupt_fn"))


def test_header_declares_the_boundary():
    names = declared_functions()
    for required in ("ipxgpu_create", "ipxgpu_normal_prepare", "ipxgpu_normal_apply",
                     "ipxgpu_diag_factorize", "ipxgpu_diag_apply", "ipxgpu_pcr_solve",
                     "ipxgpu_cr_solve", "ipxgpu_kktdiag_factorize", "ipxgpu_kktdiag_solve",
                     "ipxgpu_lu_load", "ipxgpu_tri_solve", "ipxgpu_split_prepare",
                     "ipxgpu_split_apply", "ipxgpu_comm_init"):
        assert required in names


def test_library_exports_every_declared_symbol():
    from ipx_b200 import capi
    lib = capi.load()
    missing = [n for n in declared_functions() if not hasattr(lib, n)]
    assert not missing, missing
    assert sorted(capi.EXPORTS) == declared_functions()


def test_no_cpu_fallback_without_device():
    """On a box without a GPU the product path must fail loudly, not compute."""
    from ipx_b200 import capi
    import numpy as np
    lib = capi.load()
    n = ctypes.c_int(-1)
    rc = lib.ipxgpu_device_count(ctypes.byref(n))
    if rc == 0 and n.value > 0:
        pytest.skip("a CUDA device is present")
    with pytest.raises(capi.IpxGpuError):
        capi.Context(1, 1, np.array([0, 1, 2]), np.array([0, 0]), np.array([1.0, 1.0]))


def test_product_code_never_touches_the_oracle():
    """Nothing under ipx_b200/ may import, link or open oracle/."""
    bad = []
    for root, _, files in os.walk(os.path.join(REPO, "ipx_b200")):
        if "_build" in root:
            continue
        for f in files:
            if not f.endswith((".py", ".cc", ".h", ".cu", ".cuh", ".inc")):
                continue
            text = open(os.path.join(root, f), errors="ignore").read()
            for line in text.splitlines():
                s = line.strip()
                if s.startswith(("#", "//", "*", '"""')):
                    continue
                if re.search(r"(import\s+oracle|from\s+oracle|liboracle|ipx_oracle\.h|pyoracle|"
                             r"[\"']oracle[\"']|oracle/|ipx_harness|libipx_ref)", s):
                    bad.append((f, s))
    assert not bad, bad
