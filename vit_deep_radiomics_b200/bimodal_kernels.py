"""Forward/backward of the bimodal PET+CT classifier through libvdr kernels (scope row N4).

reference: src/models_archs.py:38-124 (TransformerNoduleBimodalClassifier), :174-183 (CrossAttentionLayer), :186-200 (MLPLayer).

    y_ct  = Encoder_ct (LN(cat(cls_ct,  x_ct)))          the unimodal encoder of classifier_kernels.py, per modality
    y_pet = Encoder_pet(LN(cat(cls_pet, x_pet)))
    a_ct  = CrossAttn_ct (query = y_ct,  key = value = y_pet)[0]     only the CLS query row is kept (:102-103), so the
    a_pet = CrossAttn_pet(query = y_pet, key = value = y_ct )[0]     attention is evaluated for that one row: q0 = W_q y[0] + b_q
                                                                     (vector kernel), [K|V] = y_other W_kv^T + b_kv (tcgen05 GEMM),
                                                                     softmax / PV per head (one CTA each), out-projection (vector kernel)
    logits_ct = MLP_ct(a_ct), logits_pet = MLP_pet(a_pet)
    z = MLP_proj(cat(a_ct, a_pet)); logits_petct = MLP_petct(z)       returns (logits_petct, z, logits_ct, logits_pet)

With one modality missing the model is the unimodal classifier with that modality's encoder and head (:105-119).
Train mode: dropout 0.5 inside both encoders (:51,58) and 0.1 in the four MLPLayers (:66-75), applied inside the kernels as in
classifier_kernels.py (``drop``: a DropCfg; the two encoders and the four heads use disjoint site ids); the reference's
nn.MultiheadAttention cross attention has no dropout (:177).
"""
from __future__ import annotations

import math

import torch

from . import classifier_kernels as ck
from . import ops

_f32, _bf16, _bf16_t = ck._f32, ck._bf16, ck._bf16_t


def _head_fwd(vec_f32, hp, drop=None):
    """MLPLayer on one vector: returns (out f32, saved)."""
    v16 = vec_f32.to(torch.bfloat16)
    out, zc = ops.cls_head_fwd(v16, _f32(hp[0]), _f32(hp[1]), _f32(hp[2]), _f32(hp[3]), drop=drop)
    return out, (v16, zc, drop)


def _cross_fwd(y_q, y_kv, cp, heads):
    """cp = [in_proj_weight (3d, d), in_proj_bias (3d), out_proj.weight (d, d), out_proj.bias (d)].  Returns (a (d) f32, saved)."""
    w_in, b_in, w_o, b_o = cp
    d = y_q.shape[1]
    xq = y_q[0].float()
    q0 = ops.linear_vec_fwd(_f32(w_in)[:d], _f32(b_in)[:d], xq)
    kv = ops.gemm(y_kv, _bf16(w_in)[d:], _f32(b_in)[d:])                       # (n_kv, 2d) = [K | V]
    o, p = ops.cross_cls_attn_fwd(q0, kv, heads, 1.0 / math.sqrt(64))
    a = ops.linear_vec_fwd(_f32(w_o), _f32(b_o), o)
    return a, dict(xq=xq, q0=q0, kv=kv, p=p, o=o, y_kv=y_kv)


def _cross_bwd(da, cp, s, heads, dy_q_row0, dy_kv):
    """Accumulates into dy_q_row0 (d f32: gradient of the query modality's CLS row) and returns (param grads, dy_kv + dKV path)."""
    w_in, b_in, w_o, b_o = cp
    d = w_o.shape[0]
    dev = w_o.device
    g_in = torch.zeros(w_in.shape, dtype=torch.float32, device=dev)
    g_bin = torch.zeros(b_in.shape, dtype=torch.float32, device=dev)
    g_o = torch.zeros(w_o.shape, dtype=torch.float32, device=dev)
    g_bo = torch.zeros(b_o.shape, dtype=torch.float32, device=dev)
    do = torch.zeros(d, dtype=torch.float32, device=dev)
    ops.linear_vec_bwd(_f32(w_o), s["o"], da, g_o, g_bo, do)
    dq0, dkv = ops.cross_cls_attn_bwd(s["q0"], s["kv"], s["p"], do, heads, 1.0 / math.sqrt(64))
    ops.linear_vec_bwd(_f32(w_in)[:d], s["xq"], dq0, g_in[:d], g_bin[:d], dy_q_row0)
    ops.colsum_accum(dkv, g_bin[d:])
    g_in[d:] = ops.gemm(ops.transpose(dkv), ops.transpose(s["y_kv"]), out_dtype=torch.float32)       # dW_kv = dKV^T y_kv
    w_kv_t = ck._cached(w_in, "bf16_t_kv", lambda w: w[d:].t().to(torch.bfloat16).contiguous())      # (d, 2d): dgrad operand
    if dy_kv is not None:
        dy_kv = ops.gemm(dkv, w_kv_t, epilogue="residual", residual=dy_kv)                            # dy_kv += dKV W_kv
    else:
        dy_kv = ops.gemm(dkv, w_kv_t)
    return [g_in, g_bin, g_o, g_bo], dy_kv


#: site bases of the two encoders / indices of the four heads (ct, pet, projection, petct)
_BASE_CT, _BASE_PET = 0, 4096


def bimodal_forward(x_ct, x_pet, cfg, params, save=False, drop=None):
    """x_* (n, d) f32 CUDA or None.  cfg = dict(heads_ct, heads_pet, layers_ct, layers_pet); params = dict of parameter lists
    (enc_ct, enc_pet, cross_ct, cross_pet, head_ct, head_pet, proj, head_petct).  Returns the reference's 4-tuple (each for batch 1:
    logits (C,), petct_cls (d,), logits_ct (C,), logits_pet (C,)) [, saved]."""
    saved = {}
    y_ct = y_pet = None
    if x_ct is not None:
        r = ck.encoder_forward(x_ct, cfg["heads_ct"], cfg["layers_ct"], params["enc_ct"], save=save,
                               drop=drop.with_base(_BASE_CT) if drop is not None else None)
        y_ct, saved["enc_ct"] = r if save else (r, None)
    if x_pet is not None:
        r = ck.encoder_forward(x_pet, cfg["heads_pet"], cfg["layers_pet"], params["enc_pet"], save=save,
                               drop=drop.with_base(_BASE_PET) if drop is not None else None)
        y_pet, saved["enc_pet"] = r if save else (r, None)
    hd = (lambda i: drop.head(i)) if drop is not None else (lambda i: None)
    if y_ct is not None and y_pet is not None:
        a_ct, saved["cross_ct"] = _cross_fwd(y_ct, y_pet, params["cross_ct"], cfg["heads_ct"])
        a_pet, saved["cross_pet"] = _cross_fwd(y_pet, y_ct, params["cross_pet"], cfg["heads_ct"])   # the reference builds both with num_heads_ct (:70-71)
        logits_ct, saved["h_ct"] = _head_fwd(a_ct, params["head_ct"], hd(0))
        logits_pet, saved["h_pet"] = _head_fwd(a_pet, params["head_pet"], hd(1))
        cat = torch.cat([a_ct, a_pet])
        z, saved["h_proj"] = _head_fwd(cat, params["proj"], hd(2))
        logits, saved["h_petct"] = _head_fwd(z, params["head_petct"], hd(3))
        out = (logits, z, logits_ct, logits_pet)
        saved["mode"] = "both"
    elif y_ct is not None:
        cls = y_ct[0].float()
        logits_ct, saved["h_ct"] = _head_fwd(cls, params["head_ct"], hd(0))
        out = (logits_ct, cls, logits_ct, logits_ct)
        saved["mode"] = "ct"
    else:
        cls = y_pet[0].float()
        logits_pet, saved["h_pet"] = _head_fwd(cls, params["head_pet"], hd(1))
        out = (logits_pet, cls, logits_pet, logits_pet)
        saved["mode"] = "pet"
    return out + ((saved,) if save else ())


def bimodal_backward(cfg, params, saved, d_logits, d_z, d_logits_ct, d_logits_pet):
    """Gradients as a dict of lists mirroring `params` (None for parameter groups the forward did not use)."""
    g = {k: None for k in params}
    mode = saved["mode"]

    def add(a, b):
        if a is None:
            return b
        if b is None:
            return a
        return a + b

    if mode != "both":
        key_e, key_h, heads, layers = ("enc_ct", "head_ct", cfg["heads_ct"], cfg["layers_ct"]) if mode == "ct" else \
                                      ("enc_pet", "head_pet", cfg["heads_pet"], cfg["layers_pet"])
        dl = add(add(d_logits, d_logits_ct), d_logits_pet)            # the three logits outputs are the same tensor (:105-119)
        v16, zc, hdrop = saved["h_ct" if mode == "ct" else "h_pet"]
        g[key_h], dvec = ck.head_backward(v16, params[key_h], zc, dl, d_z, drop=hdrop)
        enc_saved = saved[key_e]
        n, d = enc_saved["x"].shape
        dy = torch.zeros((n + 1, d), dtype=torch.bfloat16, device=dvec.device)
        dy[0] = dvec.to(torch.bfloat16)
        g[key_e] = ck.encoder_backward(heads, layers, params[key_e], enc_saved, dy)
        return g

    d = saved["cross_ct"]["xq"].numel()
    v16, zc, hdrop = saved["h_petct"]
    g["head_petct"], dz = ck.head_backward(v16, params["head_petct"], zc, d_logits, d_z, drop=hdrop)
    v16, zc, hdrop = saved["h_proj"]
    g["proj"], dcat = ck.head_backward(v16, params["proj"], zc, dz, None, drop=hdrop)      # an MLPLayer whose output gradient is a vector: same kernel, C = d
    v16, zc, hdrop = saved["h_ct"]
    g["head_ct"], da_ct = ck.head_backward(v16, params["head_ct"], zc, d_logits_ct, dcat[:d].contiguous(), drop=hdrop)
    v16, zc, hdrop = saved["h_pet"]
    g["head_pet"], da_pet = ck.head_backward(v16, params["head_pet"], zc, d_logits_pet, dcat[d:].contiguous(), drop=hdrop)
    dev = da_ct.device
    row0_ct = torch.zeros(d, dtype=torch.float32, device=dev)
    row0_pet = torch.zeros(d, dtype=torch.float32, device=dev)
    g["cross_ct"], dy_pet = _cross_bwd(da_ct, params["cross_ct"], saved["cross_ct"], cfg["heads_ct"], row0_ct, None)     # keys / values: PET tokens
    g["cross_pet"], dy_ct = _cross_bwd(da_pet, params["cross_pet"], saved["cross_pet"], cfg["heads_ct"], row0_pet, None)  # keys / values: CT tokens
    dy_ct[0] = (dy_ct[0].float() + row0_ct).to(torch.bfloat16)
    dy_pet[0] = (dy_pet[0].float() + row0_pet).to(torch.bfloat16)
    g["enc_ct"] = ck.encoder_backward(cfg["heads_ct"], cfg["layers_ct"], params["enc_ct"], saved["enc_ct"], dy_ct)
    g["enc_pet"] = ck.encoder_backward(cfg["heads_pet"], cfg["layers_pet"], params["enc_pet"], saved["enc_pet"], dy_pet)
    return g


GROUPS = ("enc_ct", "enc_pet", "cross_ct", "cross_pet", "head_ct", "head_pet", "proj", "head_petct")


class BimodalFunction(torch.autograd.Function):
    """(logits_petct, petct_cls, logits_ct, logits_pet) = f(x_ct, x_pet, params...) with a hand-written backward over libvdr kernels."""

    @staticmethod
    def forward(ctx, x_ct, x_pet, cfg, sizes, drop, *flat):
        params, i = {}, 0
        for k, n in zip(GROUPS, sizes):
            params[k] = list(flat[i:i + n])
            i += n
        out = bimodal_forward(x_ct, x_pet, cfg, params, save=True, drop=drop)
        ctx.saved, ctx.params, ctx.cfg, ctx.sizes = out[4], params, cfg, sizes
        return out[0], out[1], out[2], out[3]

    @staticmethod
    def backward(ctx, d_logits, d_z, d_logits_ct, d_logits_pet):
        g = bimodal_backward(ctx.cfg, ctx.params, ctx.saved, d_logits, d_z, d_logits_ct, d_logits_pet)
        ctx.saved = None
        flat = []
        for k, n in zip(GROUPS, ctx.sizes):
            flat += list(g[k]) if g[k] is not None else [None] * n
        return (None, None, None, None, None) + tuple(flat)
