"""The C-ABI shared library loads without a GPU and exports every symbol include/vdr.h declares."""
import ctypes
import os
import re

import pytest

from vit_deep_radiomics_b200 import _C

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared_symbols():
    text = open(os.path.join(ROOT, "include", "vdr.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(vdr_[a-z0-9_]+)\s*\(", text)))


def test_header_symbols_exported():
    lib = _C.lib()
    names = declared_symbols()
    assert len(names) >= 14
    for n in names:
        assert hasattr(lib, n), f"libvdr.so does not export {n}"
    assert sorted(_C.EXPORTS) == names        # the Python binding table is in sync with the header
    assert lib.vdr_version() >= 100


def test_argument_errors_need_no_gpu():
    """Argument validation happens before any CUDA call and maps to Python exceptions."""
    lib = _C.lib()
    a = _C.GemmArgs()
    rc = lib.vdr_gemm(ctypes.byref(a), None)
    assert rc == -1 and b"null" in lib.vdr_last_error_string()
    with pytest.raises(ValueError):
        _C.check(rc, "vdr_gemm")
    assert lib.vdr_layernorm_fwd(None, 0, None, None, None, 0, 0, None, None, 1, 8, 1e-6, None) == -1
    assert lib.vdr_mask_gather_workspace_bytes(120, 32, 32, 768) >= 2 * 4 * 60 + (120 + 64) * 256 * 8


def test_missing_library_fails_loudly(monkeypatch):
    monkeypatch.setattr(_C, "_lib", None)
    monkeypatch.setattr(_C, "LIB_PATH", "/nonexistent/libvdr.so")
    with pytest.raises(_C.VdrError):
        _C.lib()


def test_ops_refuse_cpu_tensors():
    import torch
    from vit_deep_radiomics_b200 import ops
    with pytest.raises(ValueError):
        ops.gemm(torch.zeros(8, 8, dtype=torch.bfloat16), torch.zeros(8, 8, dtype=torch.bfloat16))
    with pytest.raises(ValueError):
        ops.layernorm(torch.zeros(4, 8, dtype=torch.bfloat16), torch.ones(8), torch.zeros(8), 1e-6)
