"""CPU-side checks of the CUDA product library: it builds, loads, exports every symbol the
header declares, and fails loudly (never falls back) when no GPU is present."""
import ctypes as C
import os
import re
import subprocess

import numpy as np
import pytest

from conftest import ROOT
from sac_cot_b200 import _abi


@pytest.fixture(scope="module")
def product_lib():
    path = os.path.join(ROOT, "sac_cot_b200", "lib", "libsaccot.so")
    if not os.path.exists(path):
        subprocess.run(["make", "-j", "8", "-C", os.path.join(ROOT, "sac_cot_b200", "csrc")], check=True,
                       stdout=subprocess.DEVNULL)
    from sac_cot_b200.api import load_library
    return load_library()


def header_symbols():
    text = open(os.path.join(ROOT, "include", "sac_cot.h")).read()
    return sorted(set(re.findall(r"SAC_COT_API\s+[\w\s\*]+?\b(sac_cot_\w+)\s*\(", text)))


def test_binding_table_matches_header():
    assert header_symbols() == sorted(_abi.SYMBOLS)


def test_product_library_exports_every_header_symbol(product_lib):
    for name in header_symbols():
        assert hasattr(product_lib, name), name
    assert b"cuda sm_100a" in product_lib.sac_cot_version()


def test_oracle_exports_every_header_symbol(oracle_lib):
    for name in header_symbols():
        assert hasattr(oracle_lib, name), name


def test_params_struct_layout(product_lib, oracle_lib):
    for lib in (product_lib, oracle_lib):
        p = _abi.default_params(lib)
        assert p.struct_size == C.sizeof(_abi.Params) == 40
        assert (p.num_edges, p.apex_per_edge, p.score_mode, p.refit, p.compat_mode, p.so_min_common, p.reserved) == \
            (1024, 4, 0, 1, 0, 0, 0)
        assert abs(p.tau_compat - 0.1) < 1e-7 and abs(p.tau_inlier - 0.1) < 1e-7


def test_no_cpu_fallback_without_gpu(product_lib):
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present: the no-device path cannot be exercised")
    ctx = C.c_void_p()
    rc = product_lib.sac_cot_ctx_create(C.byref(ctx), 0, None)
    assert rc == _abi.E_NODEVICE
    p = _abi.default_params(product_lib)
    src = np.zeros((16, 3), np.float32)
    R = np.zeros(9, np.float32)
    t = np.zeros(3, np.float32)
    inl = C.c_int32()
    rc = product_lib.sac_cot_register(_abi.fptr(src), _abi.fptr(src), 16, C.byref(p), _abi.fptr(R), _abi.fptr(t),
                                      C.byref(inl))
    assert rc == _abi.E_NODEVICE  # fails loudly; never computes on the CPU


def test_product_sources_never_touch_the_oracle():
    pkg = os.path.join(ROOT, "sac_cot_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".cpp", ".h")) or f == "Makefile":
                text = open(os.path.join(dirpath, f)).read()
                assert "libsaccot_oracle" not in text and "oracle/" not in text.replace("oracle/sac_cot_oracle.cpp", ""), \
                    os.path.join(dirpath, f)


def test_nccl_binding_entry_points_load_torch_first(product_lib, monkeypatch):
    """A process holds one libnccl.so.2: the Python binding makes sure torch's copy (when torch is installed) is the one
    loaded before the library's own dlopen (see _abi._prefer_torch_nccl)."""
    calls = []
    monkeypatch.setattr(_abi, "_prefer_torch_nccl", lambda: calls.append(1))
    for name in _abi._NCCL_BINDING:
        assert getattr(getattr(product_lib, name), "_sac_cot_guarded", False), name
    assert product_lib.sac_cot_ctx_set_comm(None, None, 0, 1) == _abi.E_NULL   # no ctx: refused before NCCL is touched
    assert calls == [1]
