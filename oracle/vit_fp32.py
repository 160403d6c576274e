"""TEST INFRASTRUCTURE ONLY -- plain pre-norm ViT forward in fp32/fp64 PyTorch (CPU).

Oracle for SURVEY.md section 8(a) rows V2/V2a-V2d.  PARITY UNPINNED BY THE REFERENCE: the
reference reaches its backbone through third-party packages that are not vendored and not
version-pinned (``segment_anything`` registry key 'vit_b', src/tfds_dense_descriptor.py:104,123;
``torch.hub facebookresearch/dinov2``, :87,128) and ships no golden vectors for it.  What the
reference itself fixes -- and what this file follows -- is the data flow around the backbone:

  * input  (B, 3, H, W) float32 NCHW               (prepare_image,        :45-47)
  * output (H/p, W/p, D) per slice, patch tokens only, HWC
                                                   (get_dense_descriptor, :124-133)

The arithmetic is the standard ViT of BASELINE.json's configs (ViT-S/16, B/16, L/14): conv patch
embedding + CLS token + learned absolute position embedding, L pre-norm blocks
(LN eps 1e-6, qkv bias, softmax(QK^T/sqrt(64))V, exact erf GELU, MLP ratio 4), final LN.
It is cross-checked against ``transformers.ViTModel`` with copied weights in
``tests/test_oracle_vit.py`` so that it is not merely self-consistent.

Weight names follow timm / DINOv2 (``blocks.{i}.attn.qkv.weight`` ...).
"""
from __future__ import annotations

import math

import torch
import torch.nn.functional as F

VIT_CONFIGS = {
    #            d     L   heads patch
    "vit_s16": dict(dim=384, depth=12, heads=6, patch=16),
    "vit_b16": dict(dim=768, depth=12, heads=12, patch=16),
    "vit_l14": dict(dim=1024, depth=24, heads=16, patch=14),
    # tiny config for fast CPU tests (same head_dim 64)
    "vit_t16": dict(dim=128, depth=2, heads=2, patch=16),
}


def init_weights(cfg: dict, img_hw, seed: int = 1234, dtype=torch.float32) -> dict:
    """Seeded random init: trunc_normal(std 0.02) weights, LN gamma 1 beta 0, small random biases
    (SURVEY.md section 8d asks for zero biases; small non-zero ones are used so that a dropped
    bias term is visible in parity tests)."""
    g = torch.Generator().manual_seed(seed)
    d, L, p = cfg["dim"], cfg["depth"], cfg["patch"]
    n_tok = (img_hw[0] // p) * (img_hw[1] // p) + 1

    def tn(*shape, std=0.02):
        t = torch.empty(*shape, dtype=torch.float32)
        torch.nn.init.trunc_normal_(t, std=std, a=-2 * std, b=2 * std, generator=g)
        return t.to(dtype)

    w = {
        "patch_embed.weight": tn(d, 3, p, p),
        "patch_embed.bias": tn(d, std=0.01),
        "cls_token": tn(1, 1, d),
        "pos_embed": tn(1, n_tok, d),
        "norm.weight": 1.0 + tn(d, std=0.05),
        "norm.bias": tn(d, std=0.01),
    }
    for i in range(L):
        b = f"blocks.{i}."
        w[b + "norm1.weight"] = 1.0 + tn(d, std=0.05)
        w[b + "norm1.bias"] = tn(d, std=0.01)
        w[b + "attn.qkv.weight"] = tn(3 * d, d)
        w[b + "attn.qkv.bias"] = tn(3 * d, std=0.01)
        w[b + "attn.proj.weight"] = tn(d, d)
        w[b + "attn.proj.bias"] = tn(d, std=0.01)
        w[b + "norm2.weight"] = 1.0 + tn(d, std=0.05)
        w[b + "norm2.bias"] = tn(d, std=0.01)
        w[b + "mlp.fc1.weight"] = tn(4 * d, d)
        w[b + "mlp.fc1.bias"] = tn(4 * d, std=0.01)
        w[b + "mlp.fc2.weight"] = tn(d, 4 * d)
        w[b + "mlp.fc2.bias"] = tn(d, std=0.01)
    return w


def vit_forward(w: dict, cfg: dict, x: torch.Tensor, return_tokens: bool = False):
    """x (B,3,H,W) -> dense descriptors (B, H/p, W/p, d) [and all tokens (B, N, d)].

    Layout contract: reference src/tfds_dense_descriptor.py:45-47 (NCHW in), :124-133 (HWC out).
    """
    d, L, h, p = cfg["dim"], cfg["depth"], cfg["heads"], cfg["patch"]
    dt = w["patch_embed.weight"].dtype
    x = x.to(dt)
    B, _, H, W = x.shape
    gh, gw = H // p, W // p
    t = F.conv2d(x, w["patch_embed.weight"], w["patch_embed.bias"], stride=p)     # (B,d,gh,gw)
    t = t.flatten(2).transpose(1, 2)                                              # (B,Np,d)
    t = torch.cat([w["cls_token"].expand(B, -1, -1), t], dim=1) + w["pos_embed"]  # (B,N,d)
    N = t.shape[1]
    hd = d // h
    for i in range(L):
        b = f"blocks.{i}."
        y = F.layer_norm(t, (d,), w[b + "norm1.weight"], w[b + "norm1.bias"], eps=1e-6)
        qkv = F.linear(y, w[b + "attn.qkv.weight"], w[b + "attn.qkv.bias"])
        qkv = qkv.reshape(B, N, 3, h, hd).permute(2, 0, 3, 1, 4)                  # (3,B,h,N,hd)
        q, k, v = qkv[0], qkv[1], qkv[2]
        a = torch.softmax((q @ k.transpose(-1, -2)) / math.sqrt(hd), dim=-1)
        o = (a @ v).transpose(1, 2).reshape(B, N, d)
        t = t + F.linear(o, w[b + "attn.proj.weight"], w[b + "attn.proj.bias"])
        y = F.layer_norm(t, (d,), w[b + "norm2.weight"], w[b + "norm2.bias"], eps=1e-6)
        y = F.gelu(F.linear(y, w[b + "mlp.fc1.weight"], w[b + "mlp.fc1.bias"]))    # exact erf GELU
        t = t + F.linear(y, w[b + "mlp.fc2.weight"], w[b + "mlp.fc2.bias"])
    t = F.layer_norm(t, (d,), w["norm.weight"], w["norm.bias"], eps=1e-6)
    dense = t[:, 1:, :].reshape(B, gh, gw, d)
    return (dense, t) if return_tokens else dense
