"""Forward/backward of the point-cloud classifier through libvdr kernels.

One ``torch.autograd.Function`` covers the whole network so that the saved activations and the
hand-written backward chain stay in one place.  Layer arithmetic (post-norm encoder layer of
``nn.TransformerEncoderLayer(norm_first=False, activation='gelu')``, reference
src/models_archs.py:130-147):

    y0 = LN(cat(cls, x))                                   cls_concat_layernorm kernel
    per layer:  qkv = y W_in^T + b_in                      tcgen05 GEMM (bias epilogue)
                a   = softmax(q k^T / 8) v                 fused attention kernel
                t   = a W_o^T + b_o + y                    tcgen05 GEMM (bias+residual epilogue)
                y1  = LN1(t)
                h   = gelu(y1 W_1^T + b_1)                 tcgen05 GEMM (bias+GELU epilogue)
                u   = h W_2^T + b_2 + y1                   tcgen05 GEMM (bias+residual epilogue)
                y   = LN2(u)
    cls = y[0];  logits = gelu(cls W_d1^T + b_d1) W_d2^T + b_d2
"""
from __future__ import annotations

import torch

from . import ops

#: weakly-keyed cache of bf16 operand copies: param id -> (version, tensor)
_BF16_CACHE: dict = {}


def _bf16(p: torch.Tensor, pad_rows_to: int | None = None) -> torch.Tensor:
    """bf16 copy of a weight, refreshed only when the parameter changed (optimizer step)."""
    key = (id(p), pad_rows_to)
    hit = _BF16_CACHE.get(key)
    if hit is not None and hit[0] == p._version and hit[1].device == p.device:
        return hit[1]
    w = p.detach()
    if pad_rows_to is not None and w.shape[0] < pad_rows_to:
        wp = torch.zeros((pad_rows_to,) + tuple(w.shape[1:]), dtype=w.dtype, device=w.device)
        wp[: w.shape[0]] = w
        w = wp
    w = w.to(torch.bfloat16).contiguous()
    _BF16_CACHE[key] = (p._version, w)
    return w


def _f32_padded(p: torch.Tensor, n: int) -> torch.Tensor:
    key = (id(p), "f32pad", n)
    hit = _BF16_CACHE.get(key)
    if hit is not None and hit[0] == p._version and hit[1].device == p.device:
        return hit[1]
    out = torch.zeros(n, dtype=torch.float32, device=p.device)
    out[: p.shape[0]] = p.detach()
    _BF16_CACHE[key] = (p._version, out)
    return out


def classifier_forward(x, num_heads, num_layers, params, save=False):
    """x (n, d) f32 CUDA.  Returns (logits (C,) f32, cls (d,) f32[, saved activations])."""
    n, d = x.shape
    N = n + 1
    it = iter(params)
    cls_tok, norm_w, norm_b = next(it), next(it), next(it)
    saved = {}
    if save:
        y, mu0, rs0 = ops.cls_concat_layernorm(x.contiguous(), cls_tok.detach().reshape(d).contiguous(),
                                               norm_w.detach(), norm_b.detach(), 1e-5, save_stats=True)
        saved["ln0"] = (mu0, rs0)
    else:
        y = ops.cls_concat_layernorm(x.contiguous(), cls_tok.detach().reshape(d).contiguous(), norm_w.detach(),
                                     norm_b.detach(), 1e-5)
    layers = []
    for _ in range(num_layers):
        (w_in, b_in, w_o, b_o, n1w, n1b, w1, b1, w2, b2, n2w, n2b) = (next(it) for _ in range(12))
        qkv = ops.gemm(y, _bf16(w_in), b_in.detach())
        if save:
            a, lse = ops.flash_attn(qkv, 1, N, num_heads, return_lse=True)
        else:
            a, lse = ops.flash_attn(qkv, 1, N, num_heads), None
        t = ops.gemm(a, _bf16(w_o), b_o.detach(), epilogue="residual", residual=y)
        if save:
            y1, mu1, rs1 = ops.layernorm(t, n1w.detach(), n1b.detach(), 1e-5, save_stats=True)
            z = ops.gemm(y1, _bf16(w1), b1.detach())                       # pre-activation kept for GELU'
            h = ops.gelu(z)
        else:
            y1 = ops.layernorm(t, n1w.detach(), n1b.detach(), 1e-5)
            h = ops.gemm(y1, _bf16(w1), b1.detach(), epilogue="gelu")
        u = ops.gemm(h, _bf16(w2), b2.detach(), epilogue="residual", residual=y1)
        if save:
            y2, mu2, rs2 = ops.layernorm(u, n2w.detach(), n2b.detach(), 1e-5, save_stats=True)
            layers.append(dict(y_in=y, qkv=qkv, a=a, lse=lse, t=t, mu1=mu1, rs1=rs1, y1=y1, z=z, h=h, u=u,
                               mu2=mu2, rs2=rs2))
        else:
            y2 = ops.layernorm(u, n2w.detach(), n2b.detach(), 1e-5)
        y = y2
    wd1, bd1, wd2, bd2 = next(it), next(it), next(it), next(it)
    cls_bf = y[0:1]                                                          # (1, d) bf16
    C = wd2.shape[0]
    Cp = (C + 7) // 8 * 8
    if save:
        zc = ops.gemm(cls_bf, _bf16(wd1), bd1.detach())
        hc = ops.gelu(zc)
        saved.update(zc=zc, hc=hc)
    else:
        hc = ops.gemm(cls_bf, _bf16(wd1), bd1.detach(), epilogue="gelu")
    logits = ops.gemm(hc, _bf16(wd2, pad_rows_to=Cp), _f32_padded(bd2, Cp), out_dtype=torch.float32)[0, :C]
    cls = cls_bf[0].float()
    if save:
        saved.update(layers=layers, y_last=y)
        return logits, cls, saved
    return logits, cls


class ClassifierFunction(torch.autograd.Function):
    """logits, cls = f(x, params...) with a hand-written backward over libvdr kernels."""

    @staticmethod
    def forward(ctx, x, num_heads, num_layers, *params):
        need_grad = any(ctx.needs_input_grad[3:])
        if not need_grad:
            logits, cls = classifier_forward(x, num_heads, num_layers, params, save=False)
            return logits, cls
        logits, cls, saved = classifier_forward(x, num_heads, num_layers, params, save=True)
        ctx.saved = saved
        ctx.x = x
        ctx.params = params
        ctx.num_heads, ctx.num_layers = num_heads, num_layers
        return logits, cls

    @staticmethod
    def backward(ctx, d_logits, d_cls):
        from .classifier_backward import classifier_backward
        grads = classifier_backward(ctx.x, ctx.num_heads, ctx.num_layers, ctx.params, ctx.saved, d_logits, d_cls)
        return (None, None, None) + tuple(grads)
