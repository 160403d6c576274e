"""ctypes binding of libvdr.so (C ABI declared in include/vdr.h).

There is NO fallback: if the shared library is missing or a call fails, this raises.  The
library is built in-tree by ``__graft_entry__.build()`` / ``make -C vit_deep_radiomics_b200/csrc``.
"""
from __future__ import annotations

import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("VDR_LIB") or os.path.join(_HERE, "lib", "libvdr.so")   # VDR_LIB: kernel-variant experiments only

VDR_DTYPE_BF16, VDR_DTYPE_F32 = 0, 1
EPI_BIAS, EPI_BIAS_GELU, EPI_BIAS_RESIDUAL = 0, 1, 2

#: every symbol include/vdr.h declares (checked by tests/test_abi.py)
EXPORTS = (
    "vdr_version", "vdr_last_error_string", "vdr_launch_count",
    "vdr_dropout_apply", "vdr_dropout_mask", "vdr_gemm", "vdr_vit_forward_workspace_bytes", "vdr_vit_forward", "vdr_sam_forward_workspace_bytes", "vdr_sam_forward", "vdr_patch_embed_supported", "vdr_patch_embed_gemm", "vdr_patch_embed_gemm_gray", "vdr_im2col_patches", "vdr_volume_to_slices", "vdr_volume_to_slices_resized_workspace_bytes", "vdr_volume_to_slices_resized", "vdr_volume_to_slices_cells", "vdr_im2col_gray_bf16", "vdr_write_cls_rows",
    "vdr_layernorm_fwd", "vdr_layernorm_bwd", "vdr_cls_concat_layernorm_fwd", "vdr_row_stats", "vdr_fold_layernorm",
    "vdr_flash_attn_fwd", "vdr_flash_attn_bwd_workspace_bytes", "vdr_flash_attn_bwd",
    "vdr_mask_gather_workspace_bytes", "vdr_mask_gather", "vdr_mask_gather_table", "vdr_mask_count", "vdr_exclusive_scan_i64", "vdr_debug_set_gather_trace",
    "vdr_voxel_bbox", "vdr_mask_bbox", "vdr_voxel_gather", "vdr_rotate_workspace_bytes", "vdr_flip_rotate_volume",
    "vdr_gelu_fwd", "vdr_gelu_bwd", "vdr_transpose_bf16", "vdr_colsum_bf16", "vdr_attn_delta", "vdr_attn_p_ds",
    "vdr_cls_concat_layernorm_bwd", "vdr_cls_head_fwd", "vdr_cls_head_bwd",
    "vdr_linear_vec_fwd", "vdr_linear_vec_bwd", "vdr_cross_cls_attn_fwd", "vdr_cross_cls_attn_bwd",
    "vdr_window_rows", "vdr_relpos_tables", "vdr_attn_relpos_fwd", "vdr_flash_attn_relpos_fwd", "vdr_flash_attn_relpos_fused_fwd", "vdr_attn_relpos_windows_fwd", "vdr_im2col3x3_tokens",
)


class VitBlock(C.Structure):
    """== vdr_vit_block (include/vdr.h)"""
    _fields_ = [(n, C.c_void_p) for n in ("n1w", "n1b", "qkv_w", "qkv_b", "proj_w", "proj_b",
                                          "n2w", "n2b", "fc1_w", "fc1_b", "fc2_w", "fc2_b",
                                          "qkv_wf", "qkv_bf", "qkv_cs", "fc1_wf", "fc1_bf", "fc1_cs")]


class VitWeights(C.Structure):
    """== vdr_vit_weights (include/vdr.h)"""
    _fields_ = [("dim", C.c_int), ("depth", C.c_int), ("heads", C.c_int), ("patch", C.c_int), ("H", C.c_int), ("W", C.c_int),
                ("eps", C.c_float), ("pe_w", C.c_void_p), ("pe_ldw", C.c_int64), ("pe_b", C.c_void_p), ("cls", C.c_void_p),
                ("pos", C.c_void_p), ("norm_w", C.c_void_p), ("norm_b", C.c_void_p), ("blocks", C.POINTER(VitBlock)),
                ("pe_w_gray", C.c_void_p), ("pe_gray_ldw", C.c_int64)]


class SamBlock(C.Structure):
    """== vdr_sam_block (include/vdr.h)"""
    _fields_ = [(n, C.c_void_p) for n in ("qkv_wf", "qkv_bf", "qkv_cs", "qkv_b", "proj_w", "proj_b", "fc1_wf", "fc1_bf", "fc1_cs",
                                          "fc2_w", "fc2_b", "rel_hi", "rel_lo")] + [("window", C.c_int)]


class SamWeights(C.Structure):
    """== vdr_sam_weights (include/vdr.h)"""
    _fields_ = [("dim", C.c_int), ("depth", C.c_int), ("heads", C.c_int), ("patch", C.c_int), ("H", C.c_int), ("W", C.c_int),
                ("out_chans", C.c_int), ("eps", C.c_float), ("pe_w", C.c_void_p), ("pe_ldw", C.c_int64), ("pe_w_gray", C.c_void_p),
                ("pe_gray_ldw", C.c_int64), ("pe_b", C.c_void_p), ("pos", C.c_void_p), ("neck0", C.c_void_p), ("neck1_w", C.c_void_p),
                ("neck1_b", C.c_void_p), ("neck2", C.c_void_p), ("neck3_w", C.c_void_p), ("neck3_b", C.c_void_p),
                ("blocks", C.POINTER(SamBlock))]


class Dropout(C.Structure):
    """== vdr_dropout (include/vdr.h): counter-based dropout site; thr16 = round(p * 65536), 0 = off"""
    _fields_ = [("seed", C.c_uint64), ("site", C.c_uint32), ("thr16", C.c_uint32), ("seed_offset", C.c_void_p)]


class GemmArgs(C.Structure):
    _fields_ = [
        ("A", C.c_void_p), ("lda", C.c_int64),
        ("W", C.c_void_p), ("ldw", C.c_int64),
        ("bias", C.c_void_p),
        ("R", C.c_void_p), ("ldr", C.c_int64), ("r_dtype", C.c_int),
        ("C", C.c_void_p), ("ldc", C.c_int64), ("c_dtype", C.c_int),
        ("M", C.c_int), ("N", C.c_int), ("K", C.c_int),
        ("epilogue", C.c_int),
        ("out_group", C.c_int), ("out_group_stride", C.c_int), ("out_offset", C.c_int),
        ("res_mod", C.c_int), ("res_offset", C.c_int),
        ("ln_stats", C.c_void_p), ("ln_slots", C.c_int), ("ln_eps", C.c_float), ("ln_colsum", C.c_void_p),
        ("stats_out", C.c_void_p),
        ("drop", Dropout),
    ]


class VdrError(RuntimeError):
    pass


_lib = None


def lib() -> C.CDLL:
    """Load libvdr.so once; raise (never fall back) when it is absent."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.isfile(LIB_PATH):
        raise VdrError(
            f"{LIB_PATH} not found: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
            "or `make -C vit_deep_radiomics_b200/csrc`.  There is no CPU/PyTorch fallback.")
    L = C.CDLL(LIB_PATH)
    vp, i32, i64, f32, f64, sz = C.c_void_p, C.c_int, C.c_int64, C.c_float, C.c_double, C.c_size_t
    L.vdr_version.restype = i32
    L.vdr_last_error_string.restype = C.c_char_p
    L.vdr_launch_count.restype = C.c_uint64
    L.vdr_gemm.argtypes = [C.POINTER(GemmArgs), vp]
    L.vdr_im2col_patches.argtypes = [vp, i64, i64, i64, i64, i32, i32, i32, i32, vp, vp]
    L.vdr_volume_to_slices.argtypes = [vp, i32, i32, i32, i32, i32, i32, i32, vp, vp]
    L.vdr_im2col_gray_bf16.argtypes = [vp, i32, i32, i32, i32, vp, vp]
    L.vdr_write_cls_rows.argtypes = [vp, vp, vp, i32, i32, i32, vp]
    L.vdr_layernorm_fwd.argtypes = [vp, i64, vp, vp, vp, i64, i32, vp, vp, i32, i32, f32, vp]
    L.vdr_layernorm_bwd.argtypes = [vp, i64, vp, i64, vp, vp, vp, vp, i64, vp, vp, i32, i32, vp]
    L.vdr_cls_concat_layernorm_fwd.argtypes = [vp, vp, vp, vp, vp, vp, vp, i32, i32, f32, vp]
    L.vdr_row_stats.argtypes = [vp, i64, i32, i32, vp, vp]
    L.vdr_fold_layernorm.argtypes = [vp, i64, vp, vp, vp, i32, i32, vp, i64, vp, vp, vp]
    L.vdr_vit_forward_workspace_bytes.argtypes = [C.POINTER(VitWeights), i32]
    L.vdr_vit_forward_workspace_bytes.restype = sz
    L.vdr_vit_forward.argtypes = [C.POINTER(VitWeights), vp, i32, i32, vp, i64, vp, sz, vp]
    L.vdr_sam_forward_workspace_bytes.argtypes = [C.POINTER(SamWeights), i32]
    L.vdr_sam_forward_workspace_bytes.restype = sz
    L.vdr_sam_forward.argtypes = [C.POINTER(SamWeights), vp, vp, i32, vp, i64, vp, sz, vp]
    L.vdr_volume_to_slices_resized_workspace_bytes.argtypes = [i32, i32, i32, i32, i32]
    L.vdr_volume_to_slices_resized_workspace_bytes.restype = sz
    L.vdr_volume_to_slices_resized.argtypes = [vp, i32, i32, i32, i32, i32, i32, i32, i32, i32, vp, vp, sz, vp]
    L.vdr_volume_to_slices_cells.argtypes = [vp, i32, i32, i32, i32, i32, i32, i32, i32, i32, i32, i32, vp, vp, sz, vp]
    L.vdr_patch_embed_supported.argtypes = [i32, i32, i32]
    L.vdr_patch_embed_gemm.argtypes = [vp, i32, i32, i32, i32, i32, vp, i64, vp, vp, vp, i64, i32, vp]
    L.vdr_patch_embed_gemm_gray.argtypes = [vp, i32, i32, i32, i32, vp, i64, vp, vp, vp, i64, i32, i32, vp]
    L.vdr_flash_attn_bwd_workspace_bytes.argtypes = [i32, i32, i32]
    L.vdr_flash_attn_bwd_workspace_bytes.restype = sz
    dp = C.POINTER(Dropout)
    L.vdr_dropout_apply.argtypes = [vp, i64, vp, i64, i64, i32, dp, vp]
    L.vdr_dropout_mask.argtypes = [vp, i64, i64, dp, vp]
    L.vdr_flash_attn_bwd.argtypes = [vp, i64, vp, vp, i64, vp, vp, i64, i32, i32, i32, f32, dp, vp, sz, vp]
    L.vdr_flash_attn_fwd.argtypes = [vp, i64, vp, i64, vp, i32, i32, i32, f32, dp, vp]
    L.vdr_mask_gather_workspace_bytes.argtypes = [i32, i32, i32, i32]
    L.vdr_mask_gather_workspace_bytes.restype = sz
    L.vdr_mask_gather.argtypes = [vp, i32, i64, i64, i64, i64, vp, i64, i64, i64, vp, vp, i32, i32, i32, i32,
                                  vp, vp, vp, i32, f64, vp, C.POINTER(f64), vp, sz, vp]
    L.vdr_mask_gather_table.argtypes = [vp, i32, i64, i64, i64, i64, vp, i64, i64, i64, vp, vp, i32, i32, i32, i32,
                                        vp, vp, i32, i32, vp, i64, vp, f64, vp, C.POINTER(f64), vp, sz, vp]
    L.vdr_mask_count.argtypes = [vp, i64, i64, i64, vp, vp, i32, i32, i32, vp, vp]
    L.vdr_exclusive_scan_i64.argtypes = [vp, i32, vp, vp]
    L.vdr_debug_set_gather_trace.argtypes = [vp]
    L.vdr_rotate_workspace_bytes.argtypes = [i32, i32, i32]
    L.vdr_rotate_workspace_bytes.restype = sz
    L.vdr_flip_rotate_volume.argtypes = [vp, i32, vp, i32, i32, i32, i32, i32, C.POINTER(f64), f64, f64, vp, sz, vp]
    L.vdr_voxel_bbox.argtypes = [vp, i32, i32, i32, vp, vp]
    L.vdr_mask_bbox.argtypes = [vp, i32, i32, i32, vp, vp]
    L.vdr_voxel_gather.argtypes = [vp, vp, i32, i32, i32, vp, vp, vp, vp, vp, i32, vp]
    L.vdr_gelu_fwd.argtypes = [vp, vp, i64, i32, dp, vp]
    L.vdr_gelu_bwd.argtypes = [vp, vp, vp, i64, i32, dp, vp]
    L.vdr_transpose_bf16.argtypes = [vp, i64, vp, i64, i32, i32, vp]
    L.vdr_colsum_bf16.argtypes = [vp, i64, i32, i32, vp, vp]
    L.vdr_attn_delta.argtypes = [vp, vp, i64, i32, i32, vp, vp]
    L.vdr_attn_p_ds.argtypes = [vp, vp, vp, vp, vp, vp, i32, i64, f32, vp]
    L.vdr_cls_concat_layernorm_bwd.argtypes = [vp, vp, vp, vp, vp, vp, vp, vp, vp, i32, i32, vp]
    L.vdr_cls_head_fwd.argtypes = [vp, vp, vp, vp, vp, vp, vp, i32, i32, i32, dp, vp]
    L.vdr_cls_head_bwd.argtypes = [vp, vp, vp, vp, vp, vp, vp, vp, vp, vp, vp, i32, i32, i32, dp, vp]
    L.vdr_linear_vec_fwd.argtypes = [vp, vp, vp, vp, i32, i32, vp]
    L.vdr_linear_vec_bwd.argtypes = [vp, vp, vp, vp, vp, vp, i32, i32, vp]
    L.vdr_cross_cls_attn_fwd.argtypes = [vp, vp, i64, i32, i32, f32, vp, vp, vp]
    L.vdr_cross_cls_attn_bwd.argtypes = [vp, vp, i64, vp, vp, i32, i32, f32, vp, vp, i64, vp, vp]
    L.vdr_window_rows.argtypes = [vp, i64, vp, i64, i32, i32, i32, i32, i32, i32, vp]
    L.vdr_relpos_tables.argtypes = [vp, i64, vp, vp, vp, i32, i32, i32, i32, f32, vp]
    L.vdr_flash_attn_relpos_fwd.argtypes = [vp, i64, vp, vp, i64, i32, i32, i32, f32, vp]
    L.vdr_flash_attn_relpos_fused_fwd.argtypes = [vp, i64, vp, vp, vp, i64, i32, i32, i32, f32, vp]
    L.vdr_attn_relpos_fwd.argtypes = [vp, i64, vp, vp, vp, i64, i32, i32, i32, i32, f32, vp]
    L.vdr_attn_relpos_windows_fwd.argtypes = [vp, i64, vp, vp, vp, vp, i64, i32, i32, i32, i32, i32, f32, vp]
    L.vdr_im2col3x3_tokens.argtypes = [vp, i64, vp, i64, i32, i32, i32, i32, vp]
    for name in EXPORTS:
        fn = getattr(L, name)
        if name not in ("vdr_version", "vdr_last_error_string", "vdr_launch_count",
                        "vdr_mask_gather_workspace_bytes", "vdr_vit_forward_workspace_bytes", "vdr_sam_forward_workspace_bytes",
                        "vdr_volume_to_slices_resized_workspace_bytes", "vdr_flash_attn_bwd_workspace_bytes", "vdr_rotate_workspace_bytes"):
            fn.restype = i32
    _lib = L
    return L


def check(rc: int, what: str) -> None:
    """Map the C return code to a Python exception (reference convention: plain exceptions)."""
    if rc == 0:
        return
    msg = lib().vdr_last_error_string().decode("utf-8", "replace")
    if rc in (-1, -2, -3):
        raise ValueError(f"{what}: {msg} (VDR code {rc})")
    raise VdrError(f"{what}: {msg} (code {rc})")


def launch_count() -> int:
    return int(lib().vdr_launch_count())
