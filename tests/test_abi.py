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


def test_sam_entry_points_validate_arguments_without_a_gpu():
    """The MedSAM entry points reject bad arguments before any CUDA call (no silent fallback between the attention kernels)."""
    import torch
    from vit_deep_radiomics_b200 import ops
    lib = _C.lib()
    assert lib.vdr_window_rows(None, 0, None, 0, 1, 4, 4, 2, 8, 1, None) == -1 and b"null" in lib.vdr_last_error_string()
    buf = (ctypes.c_char * 4096)()
    p = ctypes.addressof(buf)
    p += (-p) % 16
    assert lib.vdr_window_rows(p, 12, p, 12, 1, 4, 4, 2, 12, 1, None) == -1                      # d not a multiple of 8
    assert lib.vdr_attn_relpos_windows_fwd(p, 192, p, p, p, p, 64, 1, 64, 64, 20, 1, 0.125, None) == -1   # 20 x 20 windows: 400 tokens
    assert b"resident-key" in lib.vdr_last_error_string()
    assert lib.vdr_flash_attn_relpos_fwd(p, 192, p, p, 64, 1, 6, 1, 0.125, None) == -1           # Sh = 6: not a multiple of 4
    assert b"multiple of 4" in lib.vdr_last_error_string()
    assert lib.vdr_relpos_tables(p, 192, p, p, p, 1, 300, 300, 1, 1.0, None) == -1               # 90,000 tokens: beyond the index range
    assert lib.vdr_im2col3x3_tokens(p, 12, p, 108, 1, 2, 2, 12, None) == -1
    with pytest.raises(ValueError):
        ops.window_rows(torch.zeros(16, 8, dtype=torch.bfloat16), 1, 4, 4, 2, True)
    with pytest.raises(ValueError):
        ops.attn_relpos(torch.zeros(16, 192, dtype=torch.bfloat16), 1, 4, 4, 1, torch.zeros(14, 64, dtype=torch.bfloat16),
                        torch.zeros(14, 64, dtype=torch.bfloat16))


def test_sam_forward_and_gray_patch_embedding_validate_arguments_without_a_gpu():
    """vdr_sam_forward / vdr_patch_embed_gemm_gray reject bad arguments before any CUDA call; the workspace query needs no GPU."""
    lib = _C.lib()
    buf = (ctypes.c_char * 4096)()
    p = ctypes.addressof(buf)
    p += (-p) % 256
    blocks = (_C.SamBlock * 2)()
    w = _C.SamWeights(128, 2, 2, 16, 256, 256, 64, 1e-6, p, 768, p, 256, p, p, p, p, p, p, p, p, ctypes.cast(blocks, ctypes.POINTER(_C.SamBlock)))
    need = lib.vdr_sam_forward_workspace_bytes(ctypes.byref(w), 2)
    rows = 2 * 256
    assert need >= rows * 128 * 2 * (1 + 1 + 3 + 4)                      # X, Y, QKV, MLP hidden (bf16) at least
    assert lib.vdr_sam_forward(None, p, None, 1, p, 64, p, 1 << 30, None) == -1 and b"null weights" in lib.vdr_last_error_string()
    assert lib.vdr_sam_forward(ctypes.byref(w), p, p, 1, p, 64, p, 1 << 30, None) == -1           # both input forms at once
    assert b"either gray slices" in lib.vdr_last_error_string()
    assert lib.vdr_sam_forward(ctypes.byref(w), p, None, 1, p, 64, p, 16, None) != 0              # workspace too small
    assert b"workspace too small" in lib.vdr_last_error_string()
    assert lib.vdr_sam_forward(ctypes.byref(w), p, None, 1, p, 64, p, 1 << 30, None) == -1        # blocks without folded weights / tables
    assert b"folded-LayerNorm" in lib.vdr_last_error_string()
    bad = _C.SamWeights(100, 2, 2, 16, 256, 256, 64, 1e-6, p, 768, p, 256, p, p, p, p, p, p, p, p, ctypes.cast(blocks, ctypes.POINTER(_C.SamBlock)))
    assert lib.vdr_sam_forward_workspace_bytes(ctypes.byref(bad), 1) == 0                         # dim != heads x 64
    assert lib.vdr_patch_embed_gemm_gray(None, 1, 256, 256, 16, p, 256, p, p, p, 128, 128, 0, None) == -1
    assert lib.vdr_patch_embed_gemm_gray(p, 1, 224, 224, 16, p, 256, p, p, p, 128, 128, 1, None) == -1   # 14-wide grid: no TMA im2col box
    assert b"do not tile" in lib.vdr_last_error_string()
    assert lib.vdr_patch_embed_gemm_gray(p, 1, 256, 256, 16, p, 256, p, p, p, 128, 128, -1, None) == -1  # negative token offset


def test_every_entry_point_has_declared_argument_types():
    """A ctypes function without argtypes would pass 64-bit device pointers as C ints: every exported function that takes
    arguments declares them, and their count matches the header's parameter list."""
    lib = _C.lib()
    text = open(os.path.join(ROOT, "include", "vdr.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    for name in _C.EXPORTS:
        m = re.search(r"\b" + name + r"\s*\(([^;]*?)\)\s*;", text, flags=re.S)
        assert m, name
        params = [p for p in (q.strip() for q in m.group(1).split(",")) if p and p != "void"]
        fn = getattr(lib, name)
        if not params:
            assert fn.argtypes is None or len(fn.argtypes) == 0, name
        else:
            assert fn.argtypes is not None and len(fn.argtypes) == len(params), (name, len(params), fn.argtypes and len(fn.argtypes))


def test_struct_layouts_match_the_header(tmp_path):
    """sizeof / offsetof of the C structs (compiled from include/vdr.h with gcc) against the ctypes mirrors in _C.py."""
    import shutil
    import subprocess
    if shutil.which("gcc") is None:
        pytest.skip("gcc not available")
    pairs = {"vdr_gemm_args": _C.GemmArgs, "vdr_vit_block": _C.VitBlock, "vdr_vit_weights": _C.VitWeights, "vdr_dropout": _C.Dropout,
             "vdr_sam_block": _C.SamBlock, "vdr_sam_weights": _C.SamWeights}
    lines = ['#include <stdio.h>', '#include <stddef.h>', f'#include "{os.path.join(ROOT, "include", "vdr.h")}"', "int main(void) {"]
    for cname, cls in pairs.items():
        lines.append(f'printf("{cname} %zu\\n", sizeof({cname}));')
        for field, _ in cls._fields_:
            lines.append(f'printf("{cname}.{field} %zu\\n", offsetof({cname}, {field}));')
    lines += ["return 0;", "}"]
    src = tmp_path / "layout.c"
    src.write_text("\n".join(lines))
    exe = tmp_path / "layout"
    subprocess.run(["gcc", "-o", str(exe), str(src)], check=True)
    got = dict(line.split() for line in subprocess.run([str(exe)], check=True, capture_output=True, text=True).stdout.splitlines())
    for cname, cls in pairs.items():
        assert int(got[cname]) == ctypes.sizeof(cls), cname
        for field, _ in cls._fields_:
            assert int(got[f"{cname}.{field}"]) == getattr(cls, field).offset, (cname, field)


def test_hot_kernels_are_tcgen05_tma_code():
    """Static check of the shipped library (no GPU): it holds sm_100a SASS only, and the kernels of the hot path issue tcgen05.mma
    (UTCHMMA) fed by TMA (UTMALDG) with TMEM read-back (LDTM) -- a build that lost them (wrong arch, a mma.sync rewrite) fails here
    before it reaches the GPU box.  tools/sass_report.py prints the full per-kernel table (profiles/r02_sass_report.md)."""
    import shutil
    import subprocess
    import sys
    if shutil.which("cuobjdump") is None:
        pytest.skip("cuobjdump not on PATH")
    sys.path.insert(0, os.path.join(ROOT, "tools"))
    try:
        import sass_report
    finally:
        sys.path.pop(0)
    lib = os.path.join(ROOT, "vit_deep_radiomics_b200", "lib", "libvdr.so")
    archs = set(re.findall(r"arch = (sm_\w+)", subprocess.run(["cuobjdump", "-lelf", lib], capture_output=True, text=True).stdout
                           + subprocess.run(["cuobjdump", "-res-usage", lib], capture_output=True, text=True).stdout))
    assert archs == {"sm_100a"}, archs
    per_kernel = sass_report.mnemonic_counts(lib)
    by_name = {}
    for name, c in per_kernel.items():
        by_name.setdefault(name.split("<")[0], []).append(c)
    for kernel in ("vdr::gemm_tcgen05_kernel", "vdr::flash_attn_fwd_kernel", "vdr::flash_attn_fwd_v6_kernel", "vdr::flash_attn_bwd_kernel",
                   "vdr::attn_win14_tc_kernel"):
        assert kernel in by_name, f"{kernel} missing from libvdr.so"
        for c in by_name[kernel]:
            assert c["UTCHMMA"] > 0 and c["UTMALDG"] > 0 and c["LDTM"] > 0, (kernel, dict(c))
    assert any(c["2CTA"] > 0 for c in by_name["vdr::gemm_tcgen05_kernel"])          # CTA pairs (cta_group::2) in the GEMM
    assert all(c["HMMA"] == 0 for c in by_name["vdr::gemm_tcgen05_kernel"] + by_name["vdr::attn_win14_tc_kernel"])
    for kernel in ("vdr::g1_fused_kernel", "vdr::layernorm_fwd_kernel", "vdr::rot_interp_kernel"):
        assert kernel in by_name


def test_integration_guide_stub_binds_the_shipped_library():
    """The ctypes stub INTEGRATION.md shows a reference maintainer (section 2, first code block) is executed as written: every
    symbol it binds exists, and its struct mirrors have the layout of the package's own binding."""
    text = open(os.path.join(ROOT, "INTEGRATION.md")).read()
    blocks = re.findall(r"```python\n(.*?)```", text, flags=re.S)
    exec(compile(blocks[0].replace("/path/to/this/repo", ROOT), "INTEGRATION.md section 1", "exec"), {})   # the module-level imports exist
    stub = next(b for b in blocks if b.startswith("import ctypes, torch"))
    cwd = os.getcwd()
    os.chdir(ROOT)                                   # the stub loads the library by its repo-relative path
    try:
        ns: dict = {}
        exec(compile(stub, "INTEGRATION.md section 2", "exec"), ns)
    finally:
        os.chdir(cwd)
    for name in ("check", "layernorm", "linear_gelu", "image_encoder"):
        assert callable(ns[name]), name
    sam = next(b for b in blocks if "class SamBlock" in b)                        # second block: only its struct mirrors are executable
    exec(compile(re.search(r"class SamBlock.*?\n(?=lib\.vdr_sam_forward_workspace_bytes)", sam, flags=re.S).group(0), "INTEGRATION.md 2a", "exec"), ns)
    for name in ("Dropout", "GemmArgs", "VitBlock", "VitWeights", "SamBlock", "SamWeights"):
        doc, own = ns[name], getattr(_C, name)
        assert ctypes.sizeof(doc) == ctypes.sizeof(own), name
        assert [(f[0], ctypes.sizeof(f[1])) for f in doc._fields_] == [(f[0], ctypes.sizeof(f[1])) for f in own._fields_], name
    own_lib, bound = _C.lib(), 0
    for name in _C.EXPORTS:                                                       # every prototype the stub declares == the package's
        doc_args = getattr(ns["lib"], name).argtypes
        if doc_args is not None:
            own_args = getattr(own_lib, name).argtypes
            assert [ctypes.sizeof(t) for t in doc_args] == [ctypes.sizeof(t) for t in own_args], name
            bound += 1
    assert bound >= 3
    with pytest.raises(ValueError):
        ns["check"](-1, "x")
    with pytest.raises(RuntimeError):
        ns["check"](700, "x")


def test_argument_and_return_kinds_match_the_header():
    """Per parameter: pointer / integer width / float width of the ctypes prototype == the header's C type (an int64_t bound as
    c_int, or a double as c_float, still 'works' in registers and corrupts on the stack); size_t / 64-bit returns are not truncated."""
    lib = _C.lib()
    text = open(os.path.join(ROOT, "include", "vdr.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    text = re.sub(r"//[^\n]*", "", text)
    scalar = {"int": ("int", 4), "int32_t": ("int", 4), "uint32_t": ("int", 4), "unsigned": ("int", 4), "int64_t": ("int", 8),
              "uint64_t": ("int", 8), "size_t": ("int", 8), "float": ("flt", 4), "double": ("flt", 8)}

    def c_kind(param):
        if "*" in param or "vdr_stream_t" in param:
            return ("ptr", 8)
        toks = param.replace("const", "").split()
        return scalar[" ".join(toks[:-1]) if len(toks) > 1 else toks[0]]

    def py_kind(t):
        if t is ctypes.c_float:
            return ("flt", 4)
        if t is ctypes.c_double:
            return ("flt", 8)
        if t in (ctypes.c_void_p, ctypes.c_char_p) or (isinstance(t, type) and issubclass(t, (ctypes._Pointer, ctypes.Array))):
            return ("ptr", 8)
        return ("int", ctypes.sizeof(t))

    for name in _C.EXPORTS:
        m = re.search(r"([A-Za-z_0-9\* ]+?)\s*\b" + name + r"\s*\(([^;]*?)\)\s*;", text, flags=re.S)
        assert m, name
        fn = getattr(lib, name)
        params = [p for p in (q.strip() for q in m.group(2).split(",")) if p and p != "void"]
        for i, (p, t) in enumerate(zip(params, fn.argtypes or [])):
            assert c_kind(p) == py_kind(t), (name, i, p, t)
        ret = m.group(1).strip().split("\n")[-1].strip()
        if ret == "int":
            assert fn.restype is ctypes.c_int, name
        elif ret in ("size_t", "uint64_t", "int64_t"):
            assert fn.restype is not None and ctypes.sizeof(fn.restype) == 8, name
        elif ret == "const char*":
            assert fn.restype is ctypes.c_char_p, name
        else:
            assert ret == "void", (name, ret)
