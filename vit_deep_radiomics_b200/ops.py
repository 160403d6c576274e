"""Tensor-level wrappers over the C ABI (include/vdr.h).  PyTorch is used only for device
memory and the current CUDA stream; all arithmetic happens in libvdr.so.  No fallbacks."""
from __future__ import annotations

import ctypes as C
import math

import numpy as np
import torch

from . import _C

_DT = {torch.bfloat16: _C.VDR_DTYPE_BF16, torch.float32: _C.VDR_DTYPE_F32}

#: when set to a list, gemm / flash_attn append (kind, algorithmic flops, start event, end event) per launch
#: (CUDA events on the launching stream) -- used by bench.py for the roofline numbers.
PROFILE = None


class _Prof:
    def __init__(self, kind, flops, label=""):
        self.kind, self.flops, self.label = kind, flops, label

    def __enter__(self):
        if PROFILE is not None:
            self.e0 = torch.cuda.Event(enable_timing=True)
            self.e0.record()

    def __exit__(self, *exc):
        if PROFILE is not None:
            e1 = torch.cuda.Event(enable_timing=True)
            e1.record()
            PROFILE.append((self.kind, self.flops, self.e0, e1, self.label))
        return False


def _stream() -> int:
    return torch.cuda.current_stream().cuda_stream


def _req(t: torch.Tensor, dtype, name: str) -> torch.Tensor:
    if not t.is_cuda:
        raise ValueError(f"{name} must be a CUDA tensor (libvdr has no CPU path)")
    if t.device.index != torch.cuda.current_device():
        # libvdr launches on the CUDA runtime's current device / torch's current stream of that device: a tensor that lives
        # elsewhere would be touched by a kernel of the wrong device (callers: torch.cuda.set_device(...) first)
        raise ValueError(f"{name} lives on {t.device} but the current CUDA device is cuda:{torch.cuda.current_device()}; "
                         "call torch.cuda.set_device(tensor.device) before using libvdr ops")
    if t.dtype != dtype:
        raise ValueError(f"{name} must be {dtype}, got {t.dtype}")
    return t


class Drop:
    """One dropout site of a training step (include/vdr.h, vdr_dropout): ``seed`` per forward pass, ``site`` per dropout
    module, ``p`` the module's rate.  The kernels regenerate the mask from (seed, site); nothing is stored."""
    __slots__ = ("seed", "site", "thr16", "seed_offset")

    def __init__(self, seed: int, site: int, p: float, seed_offset: torch.Tensor | None = None):
        if not 0.0 <= p < 1.0:
            raise ValueError(f"dropout probability has to be in [0, 1), got {p}")
        self.seed, self.site = int(seed) & 0xFFFFFFFFFFFFFFFF, int(site) & 0xFFFFFFFF
        self.thr16 = min(65535, int(round(p * 65536)))
        # optional one-element int64 CUDA tensor the kernels add to the seed (a CUDA-graph replay bumps it on the device)
        if seed_offset is not None and (not seed_offset.is_cuda or seed_offset.dtype != torch.int64 or seed_offset.numel() != 1):
            raise ValueError("seed_offset must be a one-element int64 CUDA tensor")
        self.seed_offset = seed_offset

    @property
    def p(self) -> float:
        return self.thr16 / 65536.0

    def c(self):
        return _C.Dropout(self.seed, self.site, self.thr16, self.seed_offset.data_ptr() if self.seed_offset is not None else None)


def _dp(drop):
    """ctypes argument for an optional Drop (None / p = 0 -> NULL)."""
    return C.byref(drop.c()) if drop is not None and drop.thr16 else None


def dropout_mask(rows: int, cols: int, drop: Drop, device) -> torch.Tensor:
    """(rows, cols) uint8, 1 = kept: the mask exactly as the kernels compute it (tests, reference computations)."""
    out = torch.empty((rows, cols), dtype=torch.uint8, device=device)
    _C.check(_C.lib().vdr_dropout_mask(out.data_ptr(), rows, cols, C.byref(drop.c()), _stream()), "vdr_dropout_mask")
    return out


def dropout_apply(x: torch.Tensor, drop: Drop) -> torch.Tensor:
    """x * mask / (1 - p) for a bf16 matrix (the backward of a dropout applied inside a GEMM epilogue)."""
    _req(x, torch.bfloat16, "x")
    if x.dim() != 2 or x.stride(1) != 1:
        raise ValueError("x must be 2-D with unit inner stride")
    out = torch.empty((x.shape[0], x.shape[1]), dtype=torch.bfloat16, device=x.device)
    _C.check(_C.lib().vdr_dropout_apply(x.data_ptr(), x.stride(0), out.data_ptr(), out.stride(0), x.shape[0], x.shape[1],
                                        C.byref(drop.c()), _stream()), "vdr_dropout_apply")
    return out


def gemm(a: torch.Tensor, w: torch.Tensor, bias: torch.Tensor | None = None, *, epilogue: str = "bias",
         residual: torch.Tensor | None = None, out: torch.Tensor | None = None, out_dtype=torch.bfloat16,
         k: int | None = None, out_rows: int | None = None, out_group=(0, 0, 0), res_mod=(0, 0),
         ln_stats: torch.Tensor | None = None, ln_colsum: torch.Tensor | None = None, ln_eps: float = 1e-6,
         stats_out: torch.Tensor | None = None, drop: "Drop | None" = None) -> torch.Tensor:
    """out = epi(a @ w.T + bias [+ residual]) on tcgen05.  a (M, K[ld]) bf16, w (N, K[ld]) bf16.

    Folded LayerNorm (include/vdr.h, vdr_gemm_args): ``ln_stats`` (slots, M, 2) f32 row statistics of ``a`` + ``ln_colsum`` (N)
    with ``w`` / ``bias`` from fold_layernorm -> out = epi(LN(a) @ W.T + b); ``stats_out`` (N/64, M, 2) f32 receives the row
    statistics of what a residual GEMM writes."""
    _req(a, torch.bfloat16, "a"), _req(w, torch.bfloat16, "w")
    if a.dim() != 2 or w.dim() != 2 or a.stride(1) != 1 or w.stride(1) != 1:
        raise ValueError("a and w must be 2-D with unit inner stride")
    M, N = a.shape[0], w.shape[0]
    K = k if k is not None else a.shape[1]
    if w.shape[1] < K or a.shape[1] < K:
        raise ValueError("K larger than the operand rows")
    epi = {"bias": _C.EPI_BIAS, "gelu": _C.EPI_BIAS_GELU, "residual": _C.EPI_BIAS_RESIDUAL}[epilogue]
    if out is None:
        out = torch.empty((out_rows if out_rows is not None else M, N), dtype=out_dtype, device=a.device)
    if out.dim() != 2 or out.stride(1) != 1 or out.shape[1] != N:
        raise ValueError("out must be (rows, N) with unit inner stride")
    args = _C.GemmArgs()
    args.A, args.lda = a.data_ptr(), a.stride(0)
    args.W, args.ldw = w.data_ptr(), w.stride(0)
    args.bias = _req(bias, torch.float32, "bias").data_ptr() if bias is not None else None
    if epi == _C.EPI_BIAS_RESIDUAL:
        if residual is None:
            raise ValueError("residual epilogue needs `residual`")
        args.R, args.ldr, args.r_dtype = residual.data_ptr(), residual.stride(0), _DT[residual.dtype]
    else:
        args.R, args.ldr, args.r_dtype = None, 0, 0
    args.C, args.ldc, args.c_dtype = out.data_ptr(), out.stride(0), _DT[out.dtype]
    args.M, args.N, args.K = M, N, K
    args.epilogue = epi
    args.out_group, args.out_group_stride, args.out_offset = out_group
    args.res_mod, args.res_offset = res_mod
    if ln_stats is not None:
        _req(ln_stats, torch.float32, "ln_stats"), _req(ln_colsum, torch.float32, "ln_colsum")
        if ln_stats.dim() != 3 or ln_stats.shape[1:] != (M, 2) or not ln_stats.is_contiguous() or ln_colsum.numel() != N:
            raise ValueError("ln_stats must be a contiguous (slots, M, 2) table and ln_colsum (N)")
        args.ln_stats, args.ln_slots, args.ln_eps, args.ln_colsum = ln_stats.data_ptr(), ln_stats.shape[0], ln_eps, ln_colsum.data_ptr()
    if stats_out is not None:
        _req(stats_out, torch.float32, "stats_out")
        if not stats_out.is_contiguous() or stats_out.numel() < (N // 64) * M * 2:
            raise ValueError("stats_out must be a contiguous f32 buffer of at least (N/64, M, 2)")
        args.stats_out = stats_out.data_ptr()
    if drop is not None and drop.thr16:      # residual epilogue: out = residual + dropout(a @ w.T + bias)
        args.drop = drop.c()
    with _Prof("gemm", 2.0 * M * N * K, f"gemm M{M} N{N} K{K} {epilogue}"):
        _C.check(_C.lib().vdr_gemm(C.byref(args), _stream()), "vdr_gemm")
    return out


def row_stats(x: torch.Tensor, out: torch.Tensor | None = None) -> torch.Tensor:
    """(1, rows, 2) f32 table of (sum, sum of squares) per row of the bf16 matrix x: the one-slot form gemm(ln_stats=) reads."""
    _req(x, torch.bfloat16, "x")
    rows, d = x.shape
    if out is None:
        out = torch.empty((1, rows, 2), dtype=torch.float32, device=x.device)
    with _Prof("ln", 2.0 * rows * d, f"row_stats {rows}x{d}"):
        _C.check(_C.lib().vdr_row_stats(x.data_ptr(), x.stride(0), rows, d, out.data_ptr(), _stream()), "vdr_row_stats")
    return out


def fold_layernorm(w: torch.Tensor, bias: torch.Tensor | None, gamma: torch.Tensor, beta: torch.Tensor):
    """(W', b', colsum) of LayerNorm(gamma, beta) folded into the Linear (w bf16 (N, K), bias f32) that consumes it."""
    _req(w, torch.bfloat16, "w"), _req(gamma, torch.float32, "gamma"), _req(beta, torch.float32, "beta")
    N, K = w.shape
    wf = torch.empty_like(w)
    bf = torch.empty(N, dtype=torch.float32, device=w.device)
    cs = torch.empty(N, dtype=torch.float32, device=w.device)
    _C.check(_C.lib().vdr_fold_layernorm(w.data_ptr(), w.stride(0), bias.data_ptr() if bias is not None else None, gamma.data_ptr(),
                                         beta.data_ptr(), N, K, wf.data_ptr(), wf.stride(0), bf.data_ptr(), cs.data_ptr(), _stream()),
             "vdr_fold_layernorm")
    return wf, bf, cs


def im2col_patches(src: torch.Tensor, strides, B: int, H: int, W: int, patch: int,
                   out: torch.Tensor | None = None) -> torch.Tensor:
    """A[(b,py,px),(c,iy,ix)] bf16 from an f32 image tensor addressed by element `strides`
    (batch, channel, row, col); channel stride 0 replicates a gray image (gray2rgb)."""
    _req(src, torch.float32, "src")
    gh, gw = H // patch, W // patch
    K = 3 * patch * patch
    ldk = (K + 7) // 8 * 8
    if out is None:
        out = torch.empty((B * gh * gw, ldk), dtype=torch.bfloat16, device=src.device)
    sb, sc, sy, sx = (int(s) for s in strides)
    _C.check(_C.lib().vdr_im2col_patches(src.data_ptr(), sb, sc, sy, sx, B, H, W, patch, out.data_ptr(), _stream()),
             "vdr_im2col_patches")
    return out


def volume_to_slices(vol: torch.Tensor, crop, out: torch.Tensor | None = None, out_hw=None, cells=None) -> torch.Tensor:
    """(H, W, S) f32 volume -> (S, ch, cw) bf16 slices of the crop window (y0, y1, x0, x1); with ``out_hw`` different
    from the window size the slices are resized as prepare_image does (tfds_dense_descriptor.py:40-44: skimage resize =
    Gaussian anti-aliasing when shrinking + order-1 resampling, mirrored borders).
    ``cells = (cell_in, cell_out)``: the slices are written in the cell-padded layout of vdr_volume_to_slices_cells
    ((S, OH / cell_in * cell_out, OW / cell_in * cell_out); ``out`` must have been zeroed once: the pad pixels are never written)."""
    _req(vol, torch.float32, "vol")
    if vol.dim() != 3 or not vol.is_contiguous():
        raise ValueError("vol must be a contiguous (H, W, S) tensor")
    H, W, S = vol.shape
    y0, y1, x0, x1 = (int(v) for v in crop)
    OH, OW = (int(v) for v in out_hw) if out_hw is not None else (y1 - y0, x1 - x0)
    resized = (OH, OW) != (y1 - y0, x1 - x0)
    need = _C.lib().vdr_volume_to_slices_resized_workspace_bytes(S, y1 - y0, x1 - x0, OH, OW) if resized else 0
    if cells is not None:
        cin, cout = (int(v) for v in cells)
        if OH % cin or OW % cin or cout < cin:
            raise ValueError(f"{OH}x{OW} slices do not tile into {cin}-pixel cells")
        shape = (S, OH // cin * cout, OW // cin * cout)
        if out is None:
            out = torch.zeros(shape, dtype=torch.bfloat16, device=vol.device)
        elif tuple(out.shape) != shape or out.dtype != torch.bfloat16 or not out.is_contiguous():
            raise ValueError(f"out must be a contiguous bf16 tensor of shape {shape}")
        ws = torch.empty(max(need, 16), dtype=torch.uint8, device=vol.device)
        _C.check(_C.lib().vdr_volume_to_slices_cells(vol.data_ptr(), H, W, S, y0, x0, y1 - y0, x1 - x0, OH, OW, cin, cout, out.data_ptr(),
                                                     ws.data_ptr(), need, _stream()), "vdr_volume_to_slices_cells")
        return out
    if resized:
        if out is None:
            out = torch.empty((S, OH, OW), dtype=torch.bfloat16, device=vol.device)
        ws = torch.empty(max(need, 16), dtype=torch.uint8, device=vol.device)
        _C.check(_C.lib().vdr_volume_to_slices_resized(vol.data_ptr(), H, W, S, y0, x0, y1 - y0, x1 - x0, OH, OW, out.data_ptr(),
                                                       ws.data_ptr(), need, _stream()), "vdr_volume_to_slices_resized")
        return out
    if out is None:
        out = torch.empty((S, y1 - y0, x1 - x0), dtype=torch.bfloat16, device=vol.device)
    _C.check(_C.lib().vdr_volume_to_slices(vol.data_ptr(), H, W, S, y0, x0, y1 - y0, x1 - x0, out.data_ptr(), _stream()),
             "vdr_volume_to_slices")
    return out


def im2col_gray_bf16(slices: torch.Tensor, patch: int, out: torch.Tensor | None = None) -> torch.Tensor:
    """(B, H, W) bf16 gray slices -> A[(b,py,px),(c,iy,ix)] bf16 with the 3 channels replicated."""
    _req(slices, torch.bfloat16, "slices")
    if slices.dim() != 3 or not slices.is_contiguous():
        raise ValueError("slices must be a contiguous (B, H, W) tensor")
    B, H, W = slices.shape
    K = 3 * patch * patch
    ldk = (K + 7) // 8 * 8
    if out is None:
        out = torch.empty((B * (H // patch) * (W // patch), ldk), dtype=torch.bfloat16, device=slices.device)
    _C.check(_C.lib().vdr_im2col_gray_bf16(slices.data_ptr(), B, H, W, patch, out.data_ptr(), _stream()),
             "vdr_im2col_gray_bf16")
    return out


def write_cls_rows(cls: torch.Tensor, pos0: torch.Tensor, x: torch.Tensor, B: int, N: int, d: int) -> None:
    _req(cls, torch.float32, "cls"), _req(pos0, torch.float32, "pos0"), _req(x, torch.bfloat16, "x")
    _C.check(_C.lib().vdr_write_cls_rows(cls.data_ptr(), pos0.data_ptr(), x.data_ptr(), B, N, d, _stream()),
             "vdr_write_cls_rows")


def patch_embed_supported(H: int, W: int, patch: int) -> bool:
    """True when the geometry tiles into TMA im2col boxes (vdr_patch_embed_supported)."""
    return bool(_C.lib().vdr_patch_embed_supported(int(H), int(W), int(patch)))


def patch_embed(images: torch.Tensor, w_pe: torch.Tensor, bias: torch.Tensor, pos: torch.Tensor, patch: int,
                out: torch.Tensor, channel_summed: bool = False, token_offset: int = 1) -> torch.Tensor:
    """Patch embedding + position embedding as one TMA-fed im2col GEMM (no materialised im2col matrix).
    images (B, H, W) [gray, reused for the 3 input channels] or (B, 3, H, W), bf16 contiguous; w_pe (d, >= 3*p*p) bf16
    (k = (c, iy, ix)); bias (d) f32; pos (N, d) f32 with N = patches + 1; out (B*N, d) bf16: rows b*N + 1.. are written."""
    _req(images, torch.bfloat16, "images"), _req(w_pe, torch.bfloat16, "w_pe"), _req(pos, torch.float32, "pos")
    _req(out, torch.bfloat16, "out"), _req(bias, torch.float32, "bias")
    if not images.is_contiguous() or images.dim() not in (3, 4) or (images.dim() == 4 and images.shape[1] != 3):
        raise ValueError("images must be contiguous (B, H, W) or (B, 3, H, W)")
    B, C = images.shape[0], (1 if images.dim() == 3 else 3)
    H, W = images.shape[-2:]
    d = w_pe.shape[0]
    N = (H // patch) * (W // patch) + token_offset
    if pos.shape != (N, d) or not pos.is_contiguous() or out.shape != (B * N, d) or out.stride(1) != 1:
        raise ValueError(f"pos must be ({N}, {d}) and out ({B * N}, {d})")
    if token_offset != 1 and not channel_summed:
        raise ValueError("token_offset other than 1 needs channel_summed weights (vdr_patch_embed_gemm_gray)")
    if channel_summed:
        # gray pictures against w_pe = W_r + W_g + W_b (d, p*p): K = p*p (vdr_patch_embed_gemm_gray)
        if C != 1 or w_pe.shape[1] < patch * patch:
            raise ValueError("channel_summed weights need gray (B, H, W) pictures and w_pe (d, >= p*p)")
        with _Prof("gemm", 2.0 * B * (N - token_offset) * d * patch * patch, f"patch-embed gemm (TMA im2col, gray: channel-summed weights) M{B * (N - token_offset)} N{d} K{patch * patch}"):
            _C.check(_C.lib().vdr_patch_embed_gemm_gray(images.data_ptr(), B, H, W, patch, w_pe.data_ptr(), w_pe.stride(0),
                                                        bias.data_ptr(), pos.data_ptr(), out.data_ptr(), out.stride(0), d, int(token_offset), _stream()),
                     "vdr_patch_embed_gemm_gray")
        return out
    with _Prof("gemm", 2.0 * B * (N - 1) * d * 3 * patch * patch, f"patch-embed gemm (TMA im2col) M{B * (N - 1)} N{d} K{3 * patch * patch}"):
        _C.check(_C.lib().vdr_patch_embed_gemm(images.data_ptr(), B, C, H, W, patch, w_pe.data_ptr(), w_pe.stride(0),
                                               bias.data_ptr(), pos.data_ptr(), out.data_ptr(), out.stride(0), d, _stream()),
                 "vdr_patch_embed_gemm")
    return out


def layernorm(x: torch.Tensor, gamma: torch.Tensor, beta: torch.Tensor, eps: float, *, out_dtype=torch.bfloat16,
              out: torch.Tensor | None = None, save_stats: bool = False):
    _req(x, torch.bfloat16, "x"), _req(gamma, torch.float32, "gamma"), _req(beta, torch.float32, "beta")
    rows, d = x.shape
    if out is None:
        out = torch.empty((rows, d), dtype=out_dtype, device=x.device)
    mean = rstd = None
    if save_stats:
        mean = torch.empty(rows, dtype=torch.float32, device=x.device)
        rstd = torch.empty(rows, dtype=torch.float32, device=x.device)
    # algorithmic bytes: read the row once, write it once
    with _Prof("ln", float(rows * d * (2 + out.element_size())), f"layernorm rows{rows} d{d} -> {str(out.dtype).split('.')[-1]}"):
        _C.check(_C.lib().vdr_layernorm_fwd(x.data_ptr(), x.stride(0), gamma.data_ptr(), beta.data_ptr(),
                                            out.data_ptr(), out.stride(0), _DT[out.dtype],
                                            mean.data_ptr() if save_stats else None,
                                            rstd.data_ptr() if save_stats else None,
                                            rows, d, float(eps), _stream()), "vdr_layernorm_fwd")
    return (out, mean, rstd) if save_stats else out


def layernorm_bwd(dy, x, gamma, mean, rstd, dgamma, dbeta):
    """Returns dx (bf16); ACCUMULATES into dgamma / dbeta (f32)."""
    rows, d = x.shape
    dx = torch.empty((rows, d), dtype=torch.bfloat16, device=x.device)
    _C.check(_C.lib().vdr_layernorm_bwd(dy.data_ptr(), dy.stride(0), x.data_ptr(), x.stride(0), gamma.data_ptr(),
                                        mean.data_ptr(), rstd.data_ptr(), dx.data_ptr(), dx.stride(0),
                                        dgamma.data_ptr(), dbeta.data_ptr(), rows, d, _stream()),
             "vdr_layernorm_bwd")
    return dx


def cls_concat_layernorm(x: torch.Tensor, cls: torch.Tensor, gamma, beta, eps: float, save_stats: bool = False):
    """Y[0] = LN(cls), Y[1+i] = LN(x[i]); x (n, d) f32 -> Y (n+1, d) bf16."""
    _req(x, torch.float32, "x"), _req(cls, torch.float32, "cls")
    n, d = x.shape
    y = torch.empty((n + 1, d), dtype=torch.bfloat16, device=x.device)
    mean = rstd = None
    if save_stats:
        mean = torch.empty(n + 1, dtype=torch.float32, device=x.device)
        rstd = torch.empty(n + 1, dtype=torch.float32, device=x.device)
    _C.check(_C.lib().vdr_cls_concat_layernorm_fwd(x.data_ptr() if n else None, cls.data_ptr(), gamma.data_ptr(),
                                                   beta.data_ptr(), y.data_ptr(),
                                                   mean.data_ptr() if save_stats else None,
                                                   rstd.data_ptr() if save_stats else None,
                                                   n, d, float(eps), _stream()), "vdr_cls_concat_layernorm_fwd")
    return (y, mean, rstd) if save_stats else y


def flash_attn(qkv: torch.Tensor, B: int, N: int, heads: int, scale: float | None = None,
               out: torch.Tensor | None = None, return_lse: bool = False, drop: "Drop | None" = None):
    """qkv (B*N, 3*heads*64) bf16 -> out (B*N, heads*64) bf16.  ``drop``: attention dropout on the normalised probabilities."""
    _req(qkv, torch.bfloat16, "qkv")
    d = heads * 64
    if qkv.shape != (B * N, 3 * d):
        raise ValueError(f"qkv must be ({B * N}, {3 * d}), got {tuple(qkv.shape)}")
    if scale is None:
        scale = 1.0 / math.sqrt(64)
    if out is None:
        out = torch.empty((B * N, d), dtype=torch.bfloat16, device=qkv.device)
    lse = torch.empty((B, heads, N), dtype=torch.float32, device=qkv.device) if return_lse else None
    with _Prof("attn", 4.0 * B * heads * N * N * 64, f"attn B{B} N{N} h{heads}"):
        _C.check(_C.lib().vdr_flash_attn_fwd(qkv.data_ptr(), qkv.stride(0), out.data_ptr(), out.stride(0),
                                             lse.data_ptr() if return_lse else None, B, N, heads, float(scale), _dp(drop),
                                             _stream()), "vdr_flash_attn_fwd")
    return (out, lse) if return_lse else out


# ----------------------------------------------------------------------------- offline augmentation (flip / rotate) on the device
ROT_POLE = float.fromhex("-0x1.126145e9ecd56p-2")      # the cubic-spline pole as folded into scipy's binary (include/vdr.h)
_FLIPS = {None: 0, "None": 0, "horizontal": 1, "vertical": 2}


def rotation_xform(shape_hw, angle):
    """(m00, m01, m10, m11, off0, off1) exactly as scipy.ndimage.rotate computes them for reshape=False (same NumPy / scipy.special
    calls, so the same doubles)."""
    from scipy import special
    c, s = special.cosdg(angle), special.sindg(angle)
    rot = np.array([[c, s], [-s, c]])
    plane = np.asarray(shape_hw)
    offset = (plane - 1) / 2 - rot @ ((plane - 1) / 2)
    return (C.c_double * 6)(rot[0, 0], rot[0, 1], rot[1, 0], rot[1, 1], offset[0], offset[1])


_ROT_WS: dict = {}


def flip_rotate_volume(vol: torch.Tensor, flip=None, angle=0, *, kind: str = "image", out: torch.Tensor | None = None) -> torch.Tensor:
    """flip_image + rotate_image of the reference's augmentation loop (tfds_dense_descriptor.py:306-350, :463-466) for a whole
    device-resident volume: (H, W, S[, C]) contiguous; kind "image" (f32 -> f32 clipped to [0, 1]), "mask_bool" (a NumPy bool mask
    uploaded as uint8) or "mask_u8" (a uint8 mask; scipy rounds instead of truncating there).  Bit-identical to the host functions."""
    kinds = {"image": (0, torch.float32), "mask_bool": (1, torch.uint8), "mask_u8": (2, torch.uint8)}
    if kind not in kinds:
        raise ValueError(f"kind must be one of {sorted(kinds)}")
    code, dt = kinds[kind]
    _req(vol, dt, "vol")
    if vol.dim() not in (3, 4) or not vol.is_contiguous():
        raise ValueError("vol must be a contiguous (H, W, S) or (H, W, S, C) tensor")
    if flip not in _FLIPS:
        raise ValueError(f"flip must be None, 'horizontal' or 'vertical', got {flip!r}")
    H, W = vol.shape[0], vol.shape[1]
    planes = vol.numel() // (H * W)
    if out is None:
        out = torch.empty_like(vol)
    rotate = int(angle) % 360 != 0
    xf, zr, zc, ws, need = None, 0.0, 0.0, None, 0
    if rotate:
        xf = rotation_xform((H, W), angle)
        zr, zc = math.pow(ROT_POLE, H + 24), math.pow(ROT_POLE, W + 24)
        need = _C.lib().vdr_rotate_workspace_bytes(H, W, planes)
        key = (str(vol.device), _stream())
        ws = _ROT_WS.get(key)
        if ws is None or ws.numel() < need:
            ws = _ROT_WS[key] = torch.empty(need, dtype=torch.uint8, device=vol.device)
    _C.check(_C.lib().vdr_flip_rotate_volume(vol.data_ptr(), code, out.data_ptr(), H, W, planes, _FLIPS[flip], 1 if rotate else 0, xf, zr, zc,
                                             ws.data_ptr() if ws is not None else None, ws.numel() if ws is not None else 0, _stream()),
             "vdr_flip_rotate_volume")
    return out


# ----------------------------------------------------------------------------- gathers
def nearest_index_map(n_out: int, n_in: int) -> np.ndarray:
    """Order-0 resize source indices (skimage.transform.resize(order=0) semantics used at
    reference src/train_models.py:151): floor((o + 0.5) * (n_in / n_out) - 0.5 + 0.5), clipped."""
    ratio = np.float64(n_in) / np.float64(n_out)
    o = np.arange(n_out, dtype=np.float64)
    idx = np.floor((o + 0.5) * ratio - 0.5 + 0.5).astype(np.int64)
    return np.clip(idx, 0, n_in - 1).astype(np.int32)


_PE_DIV_CACHE: dict = {}


def _pe_div(D: int, device, scale=10000) -> torch.Tensor:
    key = (D, str(device), scale)
    if key not in _PE_DIV_CACHE:
        div = np.array([scale ** (6 * i / D) for i in range(D // 6)], dtype=np.float64)  # train_models.py:34
        if div.size == 0:
            div = np.ones(1)
        _PE_DIV_CACHE[key] = torch.from_numpy(div).to(device)
    return _PE_DIV_CACHE[key]


_GRID_MEANS_CACHE: dict = {}


def grid_means(h: int, w: int, S: int, h_orig: int, w_orig: int, res) -> tuple:
    """Means of the reference's physical grid coordinates over ALL h*w*S grid points
    (src/train_models.py:166-176), reproducing numpy's summation so the result is bit-identical.
    Depends on the geometry only: cached (it is a host pass over every grid point)."""
    key = (h, w, S, h_orig, w_orig, tuple(float(r) for r in res))
    if key not in _GRID_MEANS_CACHE:
        if len(_GRID_MEANS_CACHE) > 256:
            _GRID_MEANS_CACHE.clear()
        _GRID_MEANS_CACHE[key] = _grid_means(h, w, S, h_orig, w_orig, res)
    return _GRID_MEANS_CACHE[key]


def _grid_means(h: int, w: int, S: int, h_orig: int, w_orig: int, res) -> tuple:
    n = np.arange(h * w * S, dtype=np.int64)
    x = ((n // S) % h / w) * w_orig * res[0]
    y = ((n // (h * S)) / h) * h_orig * res[1]
    z = (n % S) * res[2]
    return float(x.mean()), float(y.mean()), float(z.mean())


_MAP_CACHE: dict = {}


def _index_map_dev(n_out: int, n_in: int, dev) -> torch.Tensor:
    key = (n_out, n_in, str(dev))
    if key not in _MAP_CACHE:
        _MAP_CACHE[key] = torch.from_numpy(nearest_index_map(n_out, n_in)).to(dev)
    return _MAP_CACHE[key]


def _mask_geometry(mask_u8: torch.Tensor, S: int, gh: int, gw: int, feat_roi, mask_roi, mask_layout: str):
    """Pixel-mask strides, ROI extents and the order-0 resize index maps of train_models.py:151 (cached device tensors)."""
    _req(mask_u8, torch.uint8, "mask")
    if not mask_u8.is_contiguous() or mask_u8.dim() != 3:
        raise ValueError("mask must be a contiguous (S, HM, WM) tensor")
    if mask_layout == "shw":
        SM, HM, WM = mask_u8.shape
        ms, mr, mc = HM * WM, WM, 1
    elif mask_layout == "hws":
        HM, WM, SM = mask_u8.shape
        ms, mr, mc = 1, WM * SM, SM
    else:
        raise ValueError("mask_layout must be 'shw' or 'hws'")
    if SM != S:
        raise ValueError(f"mask has {SM} slices, features {S}")
    r0, r1, c0, c1 = feat_roi if feat_roi is not None else (0, gh, 0, gw)
    y0, y1, x0, x1 = mask_roi if mask_roi is not None else (0, HM, 0, WM)
    h, w, hm, wm = r1 - r0, c1 - c0, y1 - y0, x1 - x0
    if min(h, w, hm, wm) <= 0 or r1 > gh or c1 > gw or y1 > HM or x1 > WM or min(r0, c0, y0, x0) < 0:
        raise ValueError("empty or out-of-range ROI")
    dev = mask_u8.device
    return dict(h=h, w=w, hm=hm, wm=wm, r0=r0, c0=c0, ms=ms, mr=mr, mc=mc, mask_ptr=mask_u8.data_ptr() + y0 * mr + x0 * mc,
                row_map=_index_map_dev(h, hm, dev), col_map=_index_map_dev(w, wm, dev))


def mask_count(mask_u8: torch.Tensor, *, grid: tuple, feat_roi: tuple | None = None, mask_roi: tuple | None = None,
               mask_layout: str = "shw", out: torch.Tensor | None = None) -> torch.Tensor:
    """Number of tokens ``mask_gather`` would select for the same mask / ROI arguments, as a device int64[1] (vdr_mask_count):
    the row count the ranks exchange before any of them emits into the shared table."""
    S, gh, gw = (int(v) for v in grid[:3])
    geo = _mask_geometry(mask_u8, S, gh, gw, feat_roi, mask_roi, mask_layout)
    if out is None:
        out = torch.empty(1, dtype=torch.int64, device=mask_u8.device)
    _req(out, torch.int64, "out")
    _C.check(_C.lib().vdr_mask_count(geo["mask_ptr"], geo["ms"], geo["mr"], geo["mc"], geo["row_map"].data_ptr(), geo["col_map"].data_ptr(),
                                     S, geo["h"], geo["w"], out.data_ptr(), _stream()), "vdr_mask_count")
    return out


def exclusive_scan_i64(counts: torch.Tensor, out: torch.Tensor | None = None) -> torch.Tensor:
    """offsets (n + 1) int64 = exclusive prefix sums of counts (n) int64 on the device (vdr_exclusive_scan_i64)."""
    _req(counts, torch.int64, "counts")
    n = counts.numel()
    if out is None:
        out = torch.empty(n + 1, dtype=torch.int64, device=counts.device)
    _req(out, torch.int64, "out")
    if not counts.is_contiguous() or not out.is_contiguous() or out.numel() < n + 1:
        raise ValueError("counts / out must be contiguous and out hold n + 1 values")
    _C.check(_C.lib().vdr_exclusive_scan_i64(counts.data_ptr(), n, out.data_ptr(), _stream()), "vdr_exclusive_scan_i64")
    return out


_GATHER_WS: dict = {}


def _gather_workspace(nbytes: int, dev) -> torch.Tensor:
    """Scratch of the gather (block totals + PE table), kept per (device, stream): calls on one stream are ordered, so the
    buffer can be reused without a fresh allocation per call."""
    key = (str(dev), _stream())
    ws = _GATHER_WS.get(key)
    if ws is None or ws.numel() < nbytes:
        ws = _GATHER_WS[key] = torch.empty(max(nbytes, 1 << 20), dtype=torch.uint8, device=dev)
    return ws


def mask_gather(feat: torch.Tensor, mask_u8: torch.Tensor, *, cap: int | None = None, pe: dict | None = None,
                grid: tuple | None = None, feat_roi: tuple | None = None, mask_roi: tuple | None = None,
                mask_layout: str = "shw", table: dict | None = None):
    """G1 (reference src/train_models.py:143-182 on device).

    feat     (S, h, w, D) dense descriptors, or the backbone's token matrix (S*slice_rows, D) together
             with ``grid=(S, gh, gw, slice_rows, first_row)`` (CLS-first layout: slice_rows = gh*gw+1,
             first_row = 1).  bf16 or f32, CUDA, unit inner stride.
    mask_u8  uint8 CUDA pixel masks, contiguous: (S, HM, WM) with mask_layout="shw", or the (HM, WM, S) volume
             mask read in place with mask_layout="hws" (no transpose pass).
    feat_roi (r0, r1, c0, c1) feature-grid window, mask_roi (y0, y1, x0, x1) pixel window: the
             extract_roi crops of tfds_dense_descriptor.py:278-279, applied by pointer arithmetic.
    pe       dict(res=(3,), noise=(3,), scale=0.25): add the 3-D positional encoding / 4 (:178-180).
    table    dict(tokens (rows, D) f32, src (rows, 3|4) int32, row_offset int64[1] device view, patient int): write into the
             slot of a shared table that starts at row ``row_offset`` (device scalar) instead of fresh tensors.
    Returns (tokens (cap, D) f32, src (cap, 3) int32 [slice,row,col in ROI coords], count int32[1]),
    all on the device; rows >= count are unspecified.  With ``table`` the first two are the table's tensors.
    """
    if feat.dtype not in _DT or not feat.is_cuda:
        raise ValueError("feat must be a bf16 or f32 CUDA tensor")
    if feat.dim() == 4:
        if not feat.is_contiguous():
            raise ValueError("dense feat must be contiguous")
        S, gh, gw, D = feat.shape
        slice_rows, first_row, ld = gh * gw, 0, D
    elif feat.dim() == 2 and grid is not None:
        S, gh, gw, slice_rows, first_row = (int(v) for v in grid)
        D, ld = feat.shape[1], feat.stride(0)
        if feat.stride(1) != 1 or feat.shape[0] < S * slice_rows:
            raise ValueError("token matrix too small / not unit inner stride")
    else:
        raise ValueError("feat must be (S,h,w,D) or a token matrix with grid=(S,gh,gw,slice_rows,first_row)")
    geo = _mask_geometry(mask_u8, S, gh, gw, feat_roi, mask_roi, mask_layout)
    h, w, hm, wm = geo["h"], geo["w"], geo["hm"], geo["wm"]
    dev = feat.device
    count = torch.empty(1, dtype=torch.int32, device=dev)
    if table is None:
        if cap is None:
            cap = S * h * w
        tokens = torch.empty((cap, D), dtype=torch.float32, device=dev)
        src = torch.empty((cap, 3), dtype=torch.int32, device=dev)
        src_cols, patient, row_off_ptr = 3, 0, None
    else:
        tokens, src = _req(table["tokens"], torch.float32, "table tokens"), _req(table["src"], torch.int32, "table src")
        if tokens.dim() != 2 or tokens.shape[1] != D or not tokens.is_contiguous() or not src.is_contiguous() or src.shape[0] != tokens.shape[0] \
                or src.shape[1] not in (3, 4):
            raise ValueError("table: tokens (rows, D) f32 and src (rows, 3|4) int32, both contiguous")
        cap, src_cols, patient = tokens.shape[0], src.shape[1], int(table.get("patient", 0))
        ro = table.get("row_offset")
        row_off_ptr = _req(ro, torch.int64, "row_offset").data_ptr() if ro is not None else None
    ws_bytes = _C.lib().vdr_mask_gather_workspace_bytes(S, h, w, D)
    ws = _gather_workspace(ws_bytes, dev)
    pe_scale, pe_div_ptr, coef = 0.0, None, None
    if pe is not None:
        res, noise = [float(v) for v in pe["res"]], [float(v) for v in pe.get("noise", (0, 0, 0))]
        mx, my, mz = grid_means(h, w, S, hm, wm, res)
        coef = (C.c_double * 11)(float(wm), float(hm), res[0], res[1], res[2], noise[0], noise[1], noise[2], mx, my, mz)
        pe_scale = float(pe.get("scale", 0.25))
        pe_div_ptr = _pe_div(D, dev).data_ptr()
    _C.check(_C.lib().vdr_mask_gather_table(feat.data_ptr(), _DT[feat.dtype], ld, slice_rows, gw, first_row + geo["r0"] * gw + geo["c0"],
                                            geo["mask_ptr"], geo["ms"], geo["mr"], geo["mc"],
                                            geo["row_map"].data_ptr(), geo["col_map"].data_ptr(), S, h, w, D,
                                            tokens.data_ptr(), src.data_ptr(), src_cols, patient, row_off_ptr, cap, count.data_ptr(),
                                            pe_scale, pe_div_ptr, coef, ws.data_ptr(), ws.numel(), _stream()),
             "vdr_mask_gather")
    return tokens, src, count


def voxel_bbox(mask_u8: torch.Tensor) -> torch.Tensor:
    """G2 pass 1: index-space bounding box (xi_min, xi_max, yi_min, yi_max, zi_min, zi_max) int32[6]."""
    _req(mask_u8, torch.uint8, "mask")
    H, W, S = mask_u8.shape
    bbox = torch.empty(6, dtype=torch.int32, device=mask_u8.device)
    _C.check(_C.lib().vdr_voxel_bbox(mask_u8.data_ptr(), H, W, S, bbox.data_ptr(), _stream()), "vdr_voxel_bbox")
    return bbox


def mask_bbox(mask_u8: torch.Tensor, out: torch.Tensor | None = None) -> torch.Tensor:
    """(col_min, col_max, row_min, row_max, slice_min, slice_max) int32[6] of an (H, W, S) uint8 mask: the bounding
    box of the union mask over slices (reference tfds_dense_descriptor.py:257-260) without a host pass."""
    _req(mask_u8, torch.uint8, "mask")
    H, W, S = mask_u8.shape
    if out is None:
        out = torch.empty(6, dtype=torch.int32, device=mask_u8.device)
    _C.check(_C.lib().vdr_mask_bbox(mask_u8.data_ptr(), H, W, S, out.data_ptr(), _stream()), "vdr_mask_bbox")
    return out


def voxel_gather(img: torch.Tensor, mask_u8: torch.Tensor, bbox: torch.Tensor, cap: int):
    """G2 pass 2: (flat int32, raw f32, mask u8, count int32[1]) of the voxels inside bbox."""
    _req(img, torch.float32, "img"), _req(mask_u8, torch.uint8, "mask")
    H, W, S = img.shape
    dev = img.device
    flat = torch.empty(cap, dtype=torch.int32, device=dev)
    raw = torch.empty(cap, dtype=torch.float32, device=dev)
    mk = torch.empty(cap, dtype=torch.uint8, device=dev)
    count = torch.empty(1, dtype=torch.int32, device=dev)
    _C.check(_C.lib().vdr_voxel_gather(img.data_ptr(), mask_u8.data_ptr(), H, W, S, bbox.data_ptr(),
                                       flat.data_ptr(), raw.data_ptr(), mk.data_ptr(), count.data_ptr(), cap,
                                       _stream()), "vdr_voxel_gather")
    return flat, raw, mk, count


# ----------------------------------------------------------------------------- training-only kernels
def gelu(z: torch.Tensor, drop: "Drop | None" = None) -> torch.Tensor:
    """dropout(gelu(z)) for a contiguous bf16 matrix (the feed-forward block's inner dropout fused into the activation)."""
    _req(z, torch.bfloat16, "z")
    if not z.is_contiguous():
        raise ValueError("z must be contiguous")
    h = torch.empty_like(z)
    _C.check(_C.lib().vdr_gelu_fwd(z.data_ptr(), h.data_ptr(), z.numel(), z.shape[-1], _dp(drop), _stream()), "vdr_gelu_fwd")
    return h


def gelu_bwd(dh: torch.Tensor, z: torch.Tensor, drop: "Drop | None" = None) -> torch.Tensor:
    if not (z.is_contiguous() and dh.is_contiguous()):
        raise ValueError("dh and z must be contiguous")
    dz = torch.empty_like(z)
    _C.check(_C.lib().vdr_gelu_bwd(dh.data_ptr(), z.data_ptr(), dz.data_ptr(), z.numel(), z.shape[-1], _dp(drop), _stream()), "vdr_gelu_bwd")
    return dz


def transpose(x: torch.Tensor) -> torch.Tensor:
    """(rows, cols) bf16 (unit inner stride, any row pitch) -> (cols, rows) view of a buffer whose pitch is
    rounded up to 8 elements, as the GEMM's TMA descriptors need."""
    _req(x, torch.bfloat16, "x")
    rows, cols = x.shape
    ld = (rows + 7) // 8 * 8
    buf = torch.empty((cols, ld), dtype=torch.bfloat16, device=x.device)
    _C.check(_C.lib().vdr_transpose_bf16(x.data_ptr(), x.stride(0), buf.data_ptr(), ld, rows, cols, _stream()),
             "vdr_transpose_bf16")
    return buf[:, :rows]


def colsum_accum(x: torch.Tensor, out: torch.Tensor) -> None:
    """out[c] += sum_r x[r, c]   (x bf16, out f32)."""
    _req(x, torch.bfloat16, "x"), _req(out, torch.float32, "out")
    rows, cols = x.shape
    _C.check(_C.lib().vdr_colsum_bf16(x.data_ptr(), x.stride(0), rows, cols, out.data_ptr(), _stream()), "vdr_colsum_bf16")


def flash_attn_bwd(qkv: torch.Tensor, o: torch.Tensor, do: torch.Tensor, lse: torch.Tensor, B: int, N: int, heads: int,
                   scale: float | None = None, drop: "Drop | None" = None) -> torch.Tensor:
    """dqkv (B*N, 3d) bf16 from the forward's packed qkv, output o, upstream gradient do (all bf16) and lse (B, heads, N) f32."""
    _req(qkv, torch.bfloat16, "qkv"), _req(o, torch.bfloat16, "o"), _req(do, torch.bfloat16, "do"), _req(lse, torch.float32, "lse")
    d = heads * 64
    if qkv.shape != (B * N, 3 * d) or o.shape != (B * N, d) or do.shape != (B * N, d) or o.stride(0) != do.stride(0):
        raise ValueError("flash_attn_bwd: qkv (B*N, 3d), o / do (B*N, d) with equal row pitch expected")
    if scale is None:
        scale = 1.0 / math.sqrt(64)
    dqkv = torch.empty((B * N, 3 * d), dtype=torch.bfloat16, device=qkv.device)
    need = _C.lib().vdr_flash_attn_bwd_workspace_bytes(B, N, heads)
    ws = torch.empty(need, dtype=torch.uint8, device=qkv.device)
    with _Prof("attn_bwd", 10.0 * B * heads * N * N * 64, f"attn-bwd B{B} N{N} h{heads}"):
        _C.check(_C.lib().vdr_flash_attn_bwd(qkv.data_ptr(), qkv.stride(0), o.data_ptr(), do.data_ptr(), o.stride(0), lse.data_ptr(),
                                             dqkv.data_ptr(), dqkv.stride(0), B, N, heads, float(scale), _dp(drop), ws.data_ptr(), need, _stream()),
                 "vdr_flash_attn_bwd")
    return dqkv


def attn_delta(dO: torch.Tensor, O: torch.Tensor, heads: int) -> torch.Tensor:
    N = O.shape[0]
    delta = torch.empty((heads, N), dtype=torch.float32, device=O.device)
    _C.check(_C.lib().vdr_attn_delta(dO.data_ptr(), O.data_ptr(), O.stride(0), N, heads, delta.data_ptr(), _stream()),
             "vdr_attn_delta")
    return delta


def attn_p_ds(S: torch.Tensor, dP: torch.Tensor, lse: torch.Tensor, delta: torch.Tensor, N: int, scale: float):
    ldp = S.shape[1]
    P = torch.empty((N, ldp), dtype=torch.bfloat16, device=S.device)
    dS = torch.empty((N, ldp), dtype=torch.bfloat16, device=S.device)
    _C.check(_C.lib().vdr_attn_p_ds(S.data_ptr(), dP.data_ptr(), lse.data_ptr(), delta.data_ptr(), P.data_ptr(),
                                    dS.data_ptr(), N, ldp, float(scale), _stream()), "vdr_attn_p_ds")
    return P, dS


def cls_concat_layernorm_bwd(dY, X, cls, gamma, mean, rstd, dgamma, dbeta, dcls):
    n, d = X.shape
    _C.check(_C.lib().vdr_cls_concat_layernorm_bwd(dY.data_ptr(), X.data_ptr() if n else None, cls.data_ptr(),
                                                   gamma.data_ptr(), mean.data_ptr(), rstd.data_ptr(), dgamma.data_ptr(),
                                                   dbeta.data_ptr(), dcls.data_ptr(), n, d, _stream()),
             "vdr_cls_concat_layernorm_bwd")


def cls_head_fwd(cls_bf16, W1, b1, W2, b2, drop: "Drop | None" = None):
    d, H1, Cn = cls_bf16.numel(), W1.shape[0], W2.shape[0]
    zc = torch.empty(H1, dtype=torch.float32, device=W1.device)
    logits = torch.empty(Cn, dtype=torch.float32, device=W1.device)
    _C.check(_C.lib().vdr_cls_head_fwd(cls_bf16.data_ptr(), W1.data_ptr(), b1.data_ptr(), W2.data_ptr(), b2.data_ptr(),
                                       zc.data_ptr(), logits.data_ptr(), d, H1, Cn, _dp(drop), _stream()), "vdr_cls_head_fwd")
    return logits, zc


def cls_head_bwd(cls_bf16, W1, W2, zc, dlogits, dcls_in, dW1, db1, dW2, db2, drop: "Drop | None" = None):
    d, H1, Cn = cls_bf16.numel(), W1.shape[0], W2.shape[0]
    dcls = torch.empty(d, dtype=torch.float32, device=W1.device)
    _C.check(_C.lib().vdr_cls_head_bwd(cls_bf16.data_ptr(), W1.data_ptr(), W2.data_ptr(), zc.data_ptr(), dlogits.data_ptr(),
                                       dcls_in.data_ptr() if dcls_in is not None else None, dW1.data_ptr(), db1.data_ptr(),
                                       dW2.data_ptr(), db2.data_ptr(), dcls.data_ptr(), d, H1, Cn, _dp(drop), _stream()),
             "vdr_cls_head_bwd")
    return dcls


# ----------------------------------------------------------------------------- bimodal classifier pieces
def linear_vec_fwd(W: torch.Tensor, b, x: torch.Tensor) -> torch.Tensor:
    """y = W x + b on one f32 vector (W (rows, cols) f32 contiguous)."""
    _req(W, torch.float32, "W"), _req(x, torch.float32, "x")
    rows, cols = W.shape
    y = torch.empty(rows, dtype=torch.float32, device=W.device)
    _C.check(_C.lib().vdr_linear_vec_fwd(W.data_ptr(), b.data_ptr() if b is not None else None, x.data_ptr(), y.data_ptr(), rows, cols,
                                         _stream()), "vdr_linear_vec_fwd")
    return y


def linear_vec_bwd(W, x, dy, dW, db, dx):
    """ACCUMULATES dW += dy x^T, db += dy (optional), dx += W^T dy."""
    rows, cols = W.shape
    _C.check(_C.lib().vdr_linear_vec_bwd(W.data_ptr(), x.data_ptr(), dy.data_ptr(), dW.data_ptr(), db.data_ptr() if db is not None else None,
                                         dx.data_ptr(), rows, cols, _stream()), "vdr_linear_vec_bwd")


def cross_cls_attn_fwd(q0: torch.Tensor, kv: torch.Tensor, heads: int, scale: float):
    """q0 (d) f32, kv (n, 2d) bf16 -> (o (d) f32, p (heads, n) f32)."""
    _req(q0, torch.float32, "q0"), _req(kv, torch.bfloat16, "kv")
    n = kv.shape[0]
    p = torch.empty((heads, n), dtype=torch.float32, device=kv.device)
    o = torch.empty(heads * 64, dtype=torch.float32, device=kv.device)
    _C.check(_C.lib().vdr_cross_cls_attn_fwd(q0.data_ptr(), kv.data_ptr(), kv.stride(0), n, heads, float(scale), p.data_ptr(), o.data_ptr(),
                                             _stream()), "vdr_cross_cls_attn_fwd")
    return o, p


def cross_cls_attn_bwd(q0, kv, p, d_o, heads: int, scale: float):
    """-> (dq0 (d) f32, dkv (n, 2d) bf16)."""
    n = kv.shape[0]
    dq0 = torch.empty(heads * 64, dtype=torch.float32, device=kv.device)
    dkv = torch.empty((n, 2 * heads * 64), dtype=torch.bfloat16, device=kv.device)
    scratch = torch.empty((heads, n), dtype=torch.float32, device=kv.device)
    _C.check(_C.lib().vdr_cross_cls_attn_bwd(q0.data_ptr(), kv.data_ptr(), kv.stride(0), p.data_ptr(), d_o.data_ptr(), n, heads, float(scale),
                                             dq0.data_ptr(), dkv.data_ptr(), dkv.stride(0), scratch.data_ptr(), _stream()),
             "vdr_cross_cls_attn_bwd")
    return dq0, dkv


# ----------------------------------------------------------------------------- MedSAM / SAM encoder pieces
def window_rows(src: torch.Tensor, B: int, H: int, W: int, ws: int, to_windows: bool, out: torch.Tensor | None = None) -> torch.Tensor:
    """window_partition (to_windows) / window_unpartition of token-major bf16 rows: (B*H*W, d) <-> (B*nwh*nww*ws*ws, d),
    zero rows for the padding (segment_anything image_encoder.py window_partition / window_unpartition)."""
    _req(src, torch.bfloat16, "src")
    d = src.shape[1]
    nwh, nww = -(-H // ws), -(-W // ws)
    rows_in = B * H * W if to_windows else B * nwh * nww * ws * ws
    rows_out = B * nwh * nww * ws * ws if to_windows else B * H * W
    if src.dim() != 2 or src.shape[0] != rows_in or src.stride(1) != 1:
        raise ValueError(f"src must be ({rows_in}, d) with unit inner stride, got {tuple(src.shape)}")
    if out is None:
        out = torch.empty((rows_out, d), dtype=torch.bfloat16, device=src.device)
    if out.shape != (rows_out, d) or out.stride(1) != 1:
        raise ValueError(f"out must be ({rows_out}, {d})")
    with _Prof("ln", float(2 * rows_out * d * 2), f"window_rows {'partition' if to_windows else 'unpartition'} {rows_out}x{d}"):
        _C.check(_C.lib().vdr_window_rows(src.data_ptr(), src.stride(0), out.data_ptr(), out.stride(0), B, H, W, ws, d,
                                          1 if to_windows else 0, _stream()), "vdr_window_rows")
    return out


def relpos_split(rel_pos_h: torch.Tensor, rel_pos_w: torch.Tensor):
    """[rel_pos_h ; rel_pos_w] (f32, 64 columns) -> (hi, lo) bf16 with hi + lo == the table to ~2^-17 relative: the operand
    form vdr_attn_relpos_fwd multiplies the queries with.  Weight preparation (once per checkpoint), on the tables' device."""
    r = torch.cat([rel_pos_h, rel_pos_w], dim=0).to(torch.float32)
    hi = r.bfloat16()
    lo = (r - hi.float()).bfloat16()
    return hi.contiguous(), lo.contiguous()


LOG2E = 1.4426950408889634


def _check_relpos_args(qkv, BW, Sh, Sw, heads, rcat_hi, rcat_lo):
    _req(qkv, torch.bfloat16, "qkv"), _req(rcat_hi, torch.bfloat16, "rcat_hi"), _req(rcat_lo, torch.bfloat16, "rcat_lo")
    N, d = Sh * Sw, heads * 64
    if qkv.shape != (BW * N, 3 * d) or qkv.stride(1) != 1:
        raise ValueError(f"qkv must be ({BW * N}, {3 * d}), got {tuple(qkv.shape)}")
    RT = 2 * Sh - 1 + 2 * Sw - 1
    if rcat_hi.shape != (RT, 64) or rcat_lo.shape != (RT, 64) or not rcat_hi.is_contiguous() or not rcat_lo.is_contiguous():
        raise ValueError(f"rcat_hi / rcat_lo must be contiguous ({RT}, 64) tables (relpos_split; interpolate first when the extent differs)")
    return N, d


def relpos_tables(qkv: torch.Tensor, BW: int, Sh: int, Sw: int, heads: int, rcat_hi: torch.Tensor, rcat_lo: torch.Tensor,
                  out_scale: float = 1.0, out: torch.Tensor | None = None) -> torch.Tensor:
    """(BW*heads*Sh*Sw, Sh+Sw) f32: out_scale * [rel_h[q, kh] | rel_w[q, kw]] of add_decomposed_rel_pos (vdr_relpos_tables)."""
    N, d = _check_relpos_args(qkv, BW, Sh, Sw, heads, rcat_hi, rcat_lo)
    rows = BW * heads * N
    if out is None:
        out = torch.empty((rows, Sh + Sw), dtype=torch.float32, device=qkv.device)
    rel = _req(out, torch.float32, "out")
    if rel.numel() < rows * (Sh + Sw) or not rel.is_contiguous():
        raise ValueError("rel scratch too small")
    with _Prof("attn", 4.0 * rows * (2 * Sh - 1 + 2 * Sw - 1) * 64, f"relpos tables BW{BW} {Sh}x{Sw} h{heads}"):
        _C.check(_C.lib().vdr_relpos_tables(qkv.data_ptr(), qkv.stride(0), rcat_hi.data_ptr(), rcat_lo.data_ptr(), rel.data_ptr(),
                                            BW, Sh, Sw, heads, float(out_scale), _stream()), "vdr_relpos_tables")
    return rel


def attn_relpos(qkv: torch.Tensor, BW: int, Sh: int, Sw: int, heads: int, rcat_hi: torch.Tensor, rcat_lo: torch.Tensor,
                scale: float | None = None, out: torch.Tensor | None = None, rel: torch.Tensor | None = None,
                kernel: str = "auto") -> torch.Tensor:
    """Attention with SAM's decomposed relative-position bias over BW images/windows of Sh x Sw tokens.
    qkv (BW*Sh*Sw, 3*heads*64) bf16; (rcat_hi, rcat_lo) = relpos_split(rel_pos_h (2*Sh-1, 64), rel_pos_w (2*Sw-1, 64))
    -> out (BW*Sh*Sw, heads*64) bf16.
    kernel: "fused" = the tcgen05 flash kernel computing its own bias terms (one launch, no table; token grids of Sh x 64,
    Sh % 4 == 0, Sh <= 64: the global-attention blocks); "tcgen05" = bias table (vdr_relpos_tables, log2 domain; `rel` = optional
    f32 scratch of BW*heads*N*(Sh+Sw)) + the same kernel reading it; "mma" = the one-launch mma.sync kernel that builds the bias on
    chip (any extent); "auto" picks "fused" where it applies."""
    N, d = _check_relpos_args(qkv, BW, Sh, Sw, heads, rcat_hi, rcat_lo)
    if scale is None:
        scale = 1.0 / math.sqrt(64)
    if out is None:
        out = torch.empty((BW * N, d), dtype=torch.bfloat16, device=qkv.device)
    if kernel == "auto":
        kernel = "fused" if (Sw == 64 and Sh % 4 == 0 and Sh <= 64) else ("tcgen05" if (Sw == 64 and Sh % 4 == 0) else "mma")
    if kernel == "fused":
        with _Prof("attn", 4.0 * BW * heads * N * N * 64 + 2.0 * BW * heads * N * 256 * 64 * 2, f"attn+relpos fused BW{BW} N{N} h{heads}"):
            _C.check(_C.lib().vdr_flash_attn_relpos_fused_fwd(qkv.data_ptr(), qkv.stride(0), rcat_hi.data_ptr(), rcat_lo.data_ptr(), out.data_ptr(),
                                                              out.stride(0), BW, Sh, heads, float(scale), _stream()), "vdr_flash_attn_relpos_fused_fwd")
        return out
    if kernel == "tcgen05":
        table = relpos_tables(qkv, BW, Sh, Sw, heads, rcat_hi, rcat_lo, out_scale=LOG2E, out=rel)
        with _Prof("attn", 4.0 * BW * heads * N * N * 64, f"attn+relpos tcgen05 BW{BW} N{N} h{heads}"):
            _C.check(_C.lib().vdr_flash_attn_relpos_fwd(qkv.data_ptr(), qkv.stride(0), table.data_ptr(), out.data_ptr(), out.stride(0),
                                                        BW, Sh, heads, float(scale), _stream()), "vdr_flash_attn_relpos_fwd")
        return out
    if kernel != "mma":
        raise ValueError(f"unknown kernel {kernel!r}")
    with _Prof("attn", 4.0 * BW * heads * N * N * 64 + 2.0 * BW * heads * N * (Sh + Sw) * 64, f"attn+relpos BW{BW} N{N} h{heads}"):
        _C.check(_C.lib().vdr_attn_relpos_fwd(qkv.data_ptr(), qkv.stride(0), rcat_hi.data_ptr(), rcat_lo.data_ptr(), out.data_ptr(),
                                              out.stride(0), BW, Sh, Sw, heads, float(scale), _stream()), "vdr_attn_relpos_fwd")
    return out


def attn_relpos_windows(qkv: torch.Tensor, qkv_bias: torch.Tensor, B: int, gh: int, gw: int, ws: int, heads: int,
                        rcat_hi: torch.Tensor, rcat_lo: torch.Tensor, scale: float | None = None, out: torch.Tensor | None = None) -> torch.Tensor:
    """A windowed SAM block's window_partition -> attention (+ rel-pos bias) -> window_unpartition in one launch, on the
    un-partitioned rows: qkv (B*gh*gw, 3*heads*64) bf16 as the qkv GEMM writes it, qkv_bias (3*heads*64) f32 = the unfolded
    Linear bias (what a zero pad token projects to), tables for ws x ws windows -> out (B*gh*gw, heads*64) bf16."""
    _req(qkv, torch.bfloat16, "qkv"), _req(qkv_bias, torch.float32, "qkv_bias"), _req(rcat_hi, torch.bfloat16, "rcat_hi"), _req(rcat_lo, torch.bfloat16, "rcat_lo")
    d = heads * 64
    if qkv.shape != (B * gh * gw, 3 * d) or qkv.stride(1) != 1 or qkv_bias.numel() != 3 * d or not qkv_bias.is_contiguous():
        raise ValueError(f"qkv must be ({B * gh * gw}, {3 * d}) and qkv_bias ({3 * d})")
    RT = 4 * ws - 2
    if rcat_hi.shape != (RT, 64) or rcat_lo.shape != (RT, 64) or not rcat_hi.is_contiguous() or not rcat_lo.is_contiguous():
        raise ValueError(f"rcat_hi / rcat_lo must be contiguous ({RT}, 64) tables (relpos_split)")
    if scale is None:
        scale = 1.0 / math.sqrt(64)
    if out is None:
        out = torch.empty((B * gh * gw, d), dtype=torch.bfloat16, device=qkv.device)
    nwin = (-(-gh // ws)) * (-(-gw // ws))
    with _Prof("attn", B * nwin * heads * (4.0 * (ws * ws) ** 2 * 64 + 2.0 * ws * ws * 2 * ws * 64), f"attn+relpos windows B{B} {gh}x{gw} ws{ws} h{heads}"):
        _C.check(_C.lib().vdr_attn_relpos_windows_fwd(qkv.data_ptr(), qkv.stride(0), qkv_bias.data_ptr(), rcat_hi.data_ptr(), rcat_lo.data_ptr(),
                                                      out.data_ptr(), out.stride(0), B, gh, gw, ws, heads, float(scale), _stream()),
                 "vdr_attn_relpos_windows_fwd")
    return out


def im2col3x3_tokens(x: torch.Tensor, B: int, H: int, W: int, out: torch.Tensor | None = None) -> torch.Tensor:
    """(B*H*W, C) bf16 token-major map -> (B*H*W, 9*C) bf16, k = (ky, kx, c), zero padding 1."""
    _req(x, torch.bfloat16, "x")
    Cc = x.shape[1]
    if x.dim() != 2 or x.shape[0] != B * H * W or x.stride(1) != 1:
        raise ValueError(f"x must be ({B * H * W}, C)")
    if out is None:
        out = torch.empty((B * H * W, 9 * Cc), dtype=torch.bfloat16, device=x.device)
    _C.check(_C.lib().vdr_im2col3x3_tokens(x.data_ptr(), x.stride(0), out.data_ptr(), out.stride(0), B, H, W, Cc, _stream()),
             "vdr_im2col3x3_tokens")
    return out
