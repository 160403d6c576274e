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
    cls = y[0];  logits = W_d2 gelu(W_d1 cls + b_d1) + b_d2        fused head kernel (fp32)

Train mode (``drop``): the reference's dropouts -- nn.TransformerEncoderLayer(dropout=p) has four per layer (attention
probabilities, the two sub-layer outputs before their residual adds, the activation inside the feed-forward block) and MLPLayer
two (hidden units, and the logits themselves) -- are applied inside the kernels from a counter-based generator
(include/vdr.h, vdr_dropout): site = base + 4 * layer + {0: attention P, 1: after out_proj, 2: after GELU, 3: after linear2};
the backward regenerates every mask from (seed, site).

Backward: dgrad / wgrad are the same tcgen05 GEMM on transposed operands (transpose kernel), bias
grads are column sums, LayerNorm / GELU have their own backward kernels, and the attention backward
recomputes P from the saved log-sum-exp with the score matrices materialised per head
(N <= ~16k tokens here: N^2 bf16 is at most a few hundred MB of the 180 GB HBM).
"""
from __future__ import annotations

import math
import weakref

import torch

from . import ops

#: cache of bf16 operand copies: (param id, tag) -> (version, tensor); refreshed when the optimizer steps
_CACHE: dict = {}


def _cached(p: torch.Tensor, tag: str, make):
    key = (id(p), tag)
    hit = _CACHE.get(key)
    # the weak reference guards against id() reuse after a model has been freed
    if hit is not None and hit[0]() is p and hit[2].device == p.device:
        if hit[1] == p._version:
            return hit[2]
        # the parameter was updated: refresh the copy IN PLACE -- a captured CUDA graph (graph_step.GraphedTrainStep) holds its address
        val = make(p.detach())
        if val.shape == hit[2].shape and val.dtype == hit[2].dtype:
            hit[2].copy_(val)
            _CACHE[key] = (hit[0], p._version, hit[2], make)
            return hit[2]
    val = make(p.detach())
    # the entry dies with its parameter (one model per fold would otherwise leak two bf16 copies of every weight)
    _CACHE[key] = (weakref.ref(p, lambda _r, k=key: _CACHE.pop(k, None)), p._version, val, make)
    return val


def refresh_cache() -> None:
    """Bring every cached operand copy up to date (in place).  The eager path does this lazily inside the forward; a replayed
    CUDA graph runs no Python, so graph_step calls it before every replay (one pass over ~50 entries)."""
    for key, hit in list(_CACHE.items()):
        p = hit[0]()
        if p is not None and hit[1] != p._version:
            _cached(p, key[1], hit[3])


def invalidate_cache() -> None:
    """Drop every cached bf16 operand copy.  The cache keys on ``Parameter._version``; writes made through ``.data`` (EMA / SWA,
    manual clamping, ``dist.broadcast(p.data, ...)``) do not bump it -- call this after such an update."""
    _CACHE.clear()


def _bf16(p):      # (out, in) bf16: forward operand
    return _cached(p, "bf16", lambda w: w.to(torch.bfloat16).contiguous())


def _bf16_t(p):    # (in, out) bf16: dgrad operand
    return _cached(p, "bf16_t", lambda w: w.t().to(torch.bfloat16).contiguous())


def _f32(p):
    return p.detach().contiguous()


class DropCfg:
    """Dropout of one forward pass: ``seed`` (fresh per pass), rate ``p`` of the encoder layers, rate ``p_head`` of the MLPLayer
    heads, ``base`` = first site id of an encoder (two encoders of the bimodal model must not share sites)."""

    def __init__(self, seed: int, p: float, p_head: float, base: int = 0, seed_offset: torch.Tensor | None = None):
        self.seed, self.p, self.p_head, self.base = int(seed), float(p), float(p_head), int(base)
        self.seed_offset = seed_offset          # device scalar added to the seed by the kernels (CUDA-graph replays; ops.Drop)

    def site(self, layer: int, k: int):
        return ops.Drop(self.seed, self.base + 4 * layer + k, self.p, self.seed_offset) if self.p > 0 else None

    def head(self, idx: int = 0):
        return ops.Drop(self.seed, 1_000_000 + idx, self.p_head, self.seed_offset) if self.p_head > 0 else None

    def with_base(self, base: int):
        return DropCfg(self.seed, self.p, self.p_head, base, self.seed_offset)


def new_seed() -> int:
    """A fresh 62-bit dropout seed from torch's default generator (so torch.manual_seed makes a run reproducible)."""
    return int(torch.randint(0, 2 ** 62, (1,), dtype=torch.int64).item())


def encoder_forward(x, num_heads, num_layers, enc_params, save=False, drop: DropCfg | None = None):
    """CLS-concat + LayerNorm + `num_layers` post-norm encoder layers.  x (n, d) f32 CUDA; enc_params = [cls_token, norm.weight,
    norm.bias] + 12 tensors per layer (PARAM_ORDER of models_archs.param_list).  Returns y (n + 1, d) bf16 (row 0 = CLS)
    [, saved activations].  ``drop``: train-mode dropout (see the module docstring)."""
    n, d = x.shape
    N = n + 1
    it = iter(enc_params)
    cls_tok, norm_w, norm_b = next(it), next(it), next(it)
    x = x.contiguous()
    cls_vec = _f32(cls_tok).reshape(d)
    saved = {}
    if save:
        y, mu0, rs0 = ops.cls_concat_layernorm(x, cls_vec, _f32(norm_w), _f32(norm_b), 1e-5, save_stats=True)
        saved["ln0"] = (mu0, rs0)
    else:
        y = ops.cls_concat_layernorm(x, cls_vec, _f32(norm_w), _f32(norm_b), 1e-5)
    layers = []
    dsite = (lambda l, k: drop.site(l, k)) if drop is not None else (lambda l, k: None)
    for l in range(num_layers):
        (w_in, b_in, w_o, b_o, n1w, n1b, w1, b1, w2, b2, n2w, n2b) = (next(it) for _ in range(12))
        qkv = ops.gemm(y, _bf16(w_in), _f32(b_in))
        if save:
            a, lse = ops.flash_attn(qkv, 1, N, num_heads, return_lse=True, drop=dsite(l, 0))
        else:
            a, lse = ops.flash_attn(qkv, 1, N, num_heads, drop=dsite(l, 0)), None
        t = ops.gemm(a, _bf16(w_o), _f32(b_o), epilogue="residual", residual=y, drop=dsite(l, 1))
        if save or dsite(l, 2) is not None:
            if save:
                y1, mu1, rs1 = ops.layernorm(t, _f32(n1w), _f32(n1b), 1e-5, save_stats=True)
            else:
                y1 = ops.layernorm(t, _f32(n1w), _f32(n1b), 1e-5)
            z = ops.gemm(y1, _bf16(w1), _f32(b1))                      # pre-activation kept for GELU'
            h = ops.gelu(z, drop=dsite(l, 2))
        else:
            y1 = ops.layernorm(t, _f32(n1w), _f32(n1b), 1e-5)
            h = ops.gemm(y1, _bf16(w1), _f32(b1), epilogue="gelu")
        u = ops.gemm(h, _bf16(w2), _f32(b2), epilogue="residual", residual=y1, drop=dsite(l, 3))
        if save:
            y2, mu2, rs2 = ops.layernorm(u, _f32(n2w), _f32(n2b), 1e-5, save_stats=True)
            layers.append(dict(y_in=y, qkv=qkv, a=a, lse=lse, t=t, mu1=mu1, rs1=rs1, y1=y1, z=z, h=h, u=u,
                               mu2=mu2, rs2=rs2))
        else:
            y2 = ops.layernorm(u, _f32(n2w), _f32(n2b), 1e-5)
        y = y2
    if save:
        saved.update(layers=layers, x=x, drop=drop)
        return y, saved
    return y


def classifier_forward(x, num_heads, num_layers, params, save=False, drop: DropCfg | None = None):
    """x (n, d) f32 CUDA.  Returns (logits (C,) f32, cls (d,) f32[, saved activations])."""
    enc = encoder_forward(x, num_heads, num_layers, params[:-4], save=save, drop=drop)
    y, saved = enc if save else (enc, None)
    wd1, bd1, wd2, bd2 = params[-4:]
    hdrop = drop.head(0) if drop is not None else None
    logits, zc = ops.cls_head_fwd(y[0], _f32(wd1), _f32(bd1), _f32(wd2), _f32(bd2), drop=hdrop)
    cls = y[0].float()
    if save:
        saved.update(y_last=y, zc=zc, head_drop=hdrop)
        return logits, cls, saved
    return logits, cls


#: use the fused tcgen05 backward (vdr_flash_attn_bwd); False = the earlier path with materialised scores (kept as a cross-check)
FUSED_ATTENTION_BACKWARD = True


def attention_backward(qkv, a, da, lse, heads, drop=None):
    """dqkv (N, 3d) bf16 from da (N, d); ``drop`` = the forward's attention-dropout site (its mask is regenerated)."""
    if FUSED_ATTENTION_BACKWARD:
        return ops.flash_attn_bwd(qkv, a, da.contiguous(), lse, 1, qkv.shape[0], heads, drop=drop)
    if drop is not None:
        raise NotImplementedError("attention dropout needs the fused backward (FUSED_ATTENTION_BACKWARD)")
    return attention_backward_materialised(qkv, a, da, lse, heads)


def attention_backward_materialised(qkv, a, da, lse, heads):
    """dqkv (N, 3d) bf16 from da (N, d): per head, S and dP are materialised in f32, P and dS in bf16."""
    N, d3 = qkv.shape
    d = d3 // 3
    Np = (N + 7) // 8 * 8
    dev = qkv.device
    scale = 1.0 / math.sqrt(64)
    dqkv = torch.empty((N, d3), dtype=torch.bfloat16, device=dev)
    delta = ops.attn_delta(da, a, heads)
    S = torch.empty((N, Np), dtype=torch.float32, device=dev)
    dP = torch.empty((N, Np), dtype=torch.float32, device=dev)
    for h in range(heads):
        q, k, v = (qkv[:, o + h * 64:o + (h + 1) * 64] for o in (0, d, 2 * d))
        do = da[:, h * 64:(h + 1) * 64]
        ops.gemm(q, k, out=S[:, :N])                                   # S = Q K^T
        ops.gemm(do, v, out=dP[:, :N])                                 # dP = dO V^T
        P, dS = ops.attn_p_ds(S, dP, lse[0, h], delta[h], N, scale)    # P = softmax, dS = P (dP - delta) / 8
        do_t = ops.transpose(do)                                       # (64, N)
        ops.gemm(ops.transpose(P[:, :N]), do_t, out=dqkv[:, 2 * d + h * 64:2 * d + (h + 1) * 64])   # dV = P^T dO
        ops.gemm(ops.transpose(dS[:, :N]), ops.transpose(q), out=dqkv[:, d + h * 64:d + (h + 1) * 64])  # dK = dS^T Q
        ops.gemm(dS[:, :N], ops.transpose(k), out=dqkv[:, h * 64:(h + 1) * 64])                      # dQ = dS K
    return dqkv


def encoder_backward(num_heads, num_layers, enc_params, saved, dy):
    """Gradients of encoder_forward's parameters (in enc_params order, f32) from dy (n + 1, d) bf16, the gradient of its
    output rows (the unimodal classifier only feeds the CLS row; the bimodal model's cross attention feeds every row)."""
    dev = enc_params[0].device
    x = saved["x"]
    n, d = x.shape
    g = [None] * len(enc_params)

    def zeros_like_param(i):
        g[i] = torch.zeros(enc_params[i].shape, dtype=torch.float32, device=dev)
        return g[i]

    drop = saved.get("drop")
    dsite = (lambda l, k: drop.site(l, k)) if drop is not None else (lambda l, k: None)
    for l in reversed(range(num_layers)):
        base = 3 + 12 * l
        (w_in, b_in, w_o, b_o, n1w, n1b, w1, b1, w2, b2, n2w, n2b) = enc_params[base:base + 12]
        s = saved["layers"][l]
        du = ops.layernorm_bwd(dy, s["u"], _f32(n2w), s["mu2"], s["rs2"], zeros_like_param(base + 10), zeros_like_param(base + 11))
        # u = y1 + dropout2(h W2^T + b2): the branch sees the masked gradient, the residual path the plain one
        g2 = ops.dropout_apply(du, dsite(l, 3)) if dsite(l, 3) is not None else du
        ops.colsum_accum(g2, zeros_like_param(base + 9))
        g[base + 8] = ops.gemm(ops.transpose(g2), ops.transpose(s["h"]), out_dtype=torch.float32)    # dW2 = g2^T h  (h = dropped activation)
        dh = ops.gemm(g2, _bf16_t(w2))                                                         # dh = g2 W2
        dz = ops.gelu_bwd(dh, s["z"], drop=dsite(l, 2))
        ops.colsum_accum(dz, zeros_like_param(base + 7))
        g[base + 6] = ops.gemm(ops.transpose(dz), ops.transpose(s["y1"]), out_dtype=torch.float32)   # dW1 = dz^T y1
        dy1 = ops.gemm(dz, _bf16_t(w1), epilogue="residual", residual=du)                      # + residual branch
        dt = ops.layernorm_bwd(dy1, s["t"], _f32(n1w), s["mu1"], s["rs1"], zeros_like_param(base + 4), zeros_like_param(base + 5))
        g1 = ops.dropout_apply(dt, dsite(l, 1)) if dsite(l, 1) is not None else dt             # t = y + dropout1(a Wo^T + bo)
        ops.colsum_accum(g1, zeros_like_param(base + 3))
        g[base + 2] = ops.gemm(ops.transpose(g1), ops.transpose(s["a"]), out_dtype=torch.float32)    # dWo = g1^T a
        da = ops.gemm(g1, _bf16_t(w_o))
        dqkv = attention_backward(s["qkv"], s["a"], da, s["lse"], num_heads, drop=dsite(l, 0))
        ops.colsum_accum(dqkv, zeros_like_param(base + 1))
        g[base] = ops.gemm(ops.transpose(dqkv), ops.transpose(s["y_in"]), out_dtype=torch.float32)   # dWin = dqkv^T y
        dy = ops.gemm(dqkv, _bf16_t(w_in), epilogue="residual", residual=dt)

    # ---- input LayerNorm + CLS token
    mu0, rs0 = saved["ln0"]
    g_cls = torch.zeros(d, dtype=torch.float32, device=dev)
    ops.cls_concat_layernorm_bwd(dy, x, _f32(enc_params[0]).reshape(d), _f32(enc_params[1]), mu0, rs0,
                                 zeros_like_param(1), zeros_like_param(2), g_cls)
    g[0] = g_cls.reshape(enc_params[0].shape)
    return g


def head_backward(y_cls_bf16, head_params, zc, d_logits, d_cls_in, drop=None):
    """MLPLayer head backward: returns (grads of [W1, b1, W2, b2], d(input vector) f32).  ``drop`` = the forward's dropout site."""
    wd1, bd1, wd2, bd2 = head_params
    dev = wd1.device
    gs = [torch.zeros(p.shape, dtype=torch.float32, device=dev) for p in head_params]
    dlog = d_logits.detach().float().contiguous() if d_logits is not None else torch.zeros(wd2.shape[0], device=dev)
    dcls_in = d_cls_in.detach().float().contiguous() if d_cls_in is not None else None
    dvec = ops.cls_head_bwd(y_cls_bf16, _f32(wd1), _f32(wd2), zc, dlog, dcls_in, gs[0], gs[1], gs[2], gs[3], drop=drop)
    return gs, dvec


def classifier_backward(num_heads, num_layers, params, saved, d_logits, d_cls):
    """Gradients for every parameter, in ``params`` order (f32)."""
    dev = params[0].device
    n, d = saved["x"].shape
    y_last = saved["y_last"]
    g_head, dcls = head_backward(y_last[0], params[-4:], saved["zc"], d_logits, d_cls, drop=saved.get("head_drop"))
    dy = torch.zeros((n + 1, d), dtype=torch.bfloat16, device=dev)          # only the CLS row carries gradient
    dy[0] = dcls.to(torch.bfloat16)
    return encoder_backward(num_heads, num_layers, params[:-4], saved, dy) + g_head


class ClassifierFunction(torch.autograd.Function):
    """logits, cls = f(x, params...) with a hand-written backward over libvdr kernels."""

    @staticmethod
    def forward(ctx, x, num_heads, num_layers, drop, *params):
        logits, cls, saved = classifier_forward(x, num_heads, num_layers, params, save=True, drop=drop)
        ctx.saved = saved
        ctx.params = params
        ctx.num_heads, ctx.num_layers = num_heads, num_layers
        return logits, cls

    @staticmethod
    def backward(ctx, d_logits, d_cls):
        grads = classifier_backward(ctx.num_heads, ctx.num_layers, ctx.params, ctx.saved, d_logits, d_cls)
        ctx.saved = None
        return (None, None, None, None) + tuple(grads)
