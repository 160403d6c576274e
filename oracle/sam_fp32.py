"""TEST INFRASTRUCTURE ONLY (never imported by the product package).

CPU fp32 restatement of the two backbones the reference actually loads:

* ``sam_image_encoder``: ``model.image_encoder(img_tensor)`` of ``sam_model_registry['vit_b']``
  (reference call sites src/tfds_dense_descriptor.py:104 and :123).  The code lives in the third-party
  package ``segment_anything`` (facebookresearch/segment-anything, ``modeling/image_encoder.py``), which the
  reference neither vendors nor pins (no requirements file) and which is absent from this image.  The
  published algorithm is restated here: 16x16 patch embedding, absolute position embedding (1, 64, 64, d),
  pre-norm blocks with 14x14 windowed attention (zero padding 64 -> 70 applied AFTER norm1, pad tokens take
  part in the window's softmax) except at the global-attention blocks (2, 5, 8, 11), decomposed relative
  position bias computed from the UNSCALED queries, MLP with erf GELU, neck = 1x1 conv -> LayerNorm2d ->
  3x3 conv (padding 1) -> LayerNorm2d, all LayerNorm eps 1e-6.  State-dict keys are segment_anything's
  (``pos_embed``, ``patch_embed.proj.*``, ``blocks.i.{norm1,attn.qkv,attn.proj,attn.rel_pos_h,attn.rel_pos_w,
  norm2,mlp.lin1,mlp.lin2}``, ``neck.{0,1,2,3}``) so a MedSAM checkpoint loads unchanged.
  **Parity unpinned by the reference** (no golden vectors, dependency un-versioned); the restatement is pinned
  against ``transformers.models.sam.modeling_sam.SamVisionEncoder`` (same architecture, independent
  implementation) with copied weights in tests/test_oracle_sam.py.

* ``dinov2_patch_embed``: ``model.patch_embed(img_tensor)`` of the torch.hub DINOv2 ViT-S/14
  (src/tfds_dense_descriptor.py:87,128): only the 14x14 strided convolution + bias, tokens (Np, 384).
"""
from __future__ import annotations

import torch
import torch.nn.functional as F

SAM_CONFIGS = {
    # name: dim, depth, heads, global-attention blocks, window, neck channels, patch
    "medsam": dict(dim=768, depth=12, heads=12, global_attn=(2, 5, 8, 11), window=14, out_chans=256, patch=16),
    "sam_tiny": dict(dim=128, depth=4, heads=2, global_attn=(1, 3), window=14, out_chans=64, patch=16),   # tests
    "sam_small": dict(dim=128, depth=2, heads=2, global_attn=(1,), window=14, out_chans=256, patch=16),   # tests: MedSAM's 256-wide neck
}


def init_sam_state_dict(cfg: dict, img_hw, seed: int = 1234) -> dict:
    """Seeded random init with segment_anything key names (no checkpoints are available offline)."""
    g = torch.Generator().manual_seed(seed)
    d, p, oc, ws = cfg["dim"], cfg["patch"], cfg["out_chans"], cfg["window"]
    gh, gw = img_hw[0] // p, img_hw[1] // p
    hd = d // cfg["heads"]

    def tn(*shape, std=0.02):
        t = torch.empty(*shape, dtype=torch.float32)
        torch.nn.init.trunc_normal_(t, std=std, a=-2 * std, b=2 * std, generator=g)
        return t

    w = {"pos_embed": tn(1, gh, gw, d), "patch_embed.proj.weight": tn(d, 3, p, p), "patch_embed.proj.bias": tn(d, std=0.01)}
    for i in range(cfg["depth"]):
        b = f"blocks.{i}."
        sh, sw = (gh, gw) if i in cfg["global_attn"] else (ws, ws)
        w[b + "norm1.weight"] = 1.0 + tn(d, std=0.05)
        w[b + "norm1.bias"] = tn(d, std=0.01)
        w[b + "attn.qkv.weight"] = tn(3 * d, d, std=0.04)
        w[b + "attn.qkv.bias"] = tn(3 * d, std=0.01)
        w[b + "attn.proj.weight"] = tn(d, d)
        w[b + "attn.proj.bias"] = tn(d, std=0.01)
        w[b + "attn.rel_pos_h"] = tn(2 * sh - 1, hd, std=0.1)
        w[b + "attn.rel_pos_w"] = tn(2 * sw - 1, hd, std=0.1)
        w[b + "norm2.weight"] = 1.0 + tn(d, std=0.05)
        w[b + "norm2.bias"] = tn(d, std=0.01)
        w[b + "mlp.lin1.weight"] = tn(4 * d, d)
        w[b + "mlp.lin1.bias"] = tn(4 * d, std=0.01)
        w[b + "mlp.lin2.weight"] = tn(d, 4 * d)
        w[b + "mlp.lin2.bias"] = tn(d, std=0.01)
    w["neck.0.weight"] = tn(oc, d, 1, 1, std=0.05)
    w["neck.1.weight"] = 1.0 + tn(oc, std=0.05)
    w["neck.1.bias"] = tn(oc, std=0.01)
    w["neck.2.weight"] = tn(oc, oc, 3, 3, std=0.05)
    w["neck.3.weight"] = 1.0 + tn(oc, std=0.05)
    w["neck.3.bias"] = tn(oc, std=0.01)
    return w


def rel_pos_rows(size: int, rel_pos: torch.Tensor) -> torch.Tensor:
    """R[q, k, :] = rel_pos[q - k + size - 1] for a square q/k extent (segment_anything get_rel_pos with
    q_size == k_size; a table of another length is linearly interpolated to 2*size - 1 first)."""
    L = 2 * size - 1
    if rel_pos.shape[0] != L:
        rel_pos = F.interpolate(rel_pos.reshape(1, rel_pos.shape[0], -1).permute(0, 2, 1), size=L, mode="linear")
        rel_pos = rel_pos.reshape(-1, L).permute(1, 0)
    idx = torch.arange(size)[:, None] - torch.arange(size)[None, :] + (size - 1)
    return rel_pos[idx]


def attention_relpos(x: torch.Tensor, w: dict, b: str, heads: int) -> torch.Tensor:
    """x (B, H, W, d) -> (B, H, W, d): multi-head attention over the H*W tokens with the decomposed
    relative-position bias rel_h[q, kh] + rel_w[q, kw] built from the unscaled queries."""
    B, H, W, d = x.shape
    hd = d // heads
    qkv = F.linear(x.reshape(B, H * W, d), w[b + "attn.qkv.weight"], w[b + "attn.qkv.bias"])
    q, k, v = qkv.reshape(B, H * W, 3, heads, hd).permute(2, 0, 3, 1, 4).reshape(3, B * heads, H * W, hd).unbind(0)
    attn = (q * hd ** -0.5) @ k.transpose(-2, -1)
    Rh, Rw = rel_pos_rows(H, w[b + "attn.rel_pos_h"]), rel_pos_rows(W, w[b + "attn.rel_pos_w"])
    rq = q.reshape(B * heads, H, W, hd)
    rel_h = torch.einsum("bhwc,hkc->bhwk", rq, Rh)
    rel_w = torch.einsum("bhwc,wkc->bhwk", rq, Rw)
    attn = (attn.view(B * heads, H, W, H, W) + rel_h[:, :, :, :, None] + rel_w[:, :, :, None, :]).view(B * heads, H * W, H * W)
    attn = attn.softmax(dim=-1)
    out = (attn @ v).view(B, heads, H, W, hd).permute(0, 2, 3, 1, 4).reshape(B, H, W, d)
    return F.linear(out, w[b + "attn.proj.weight"], w[b + "attn.proj.bias"])


def window_partition(x: torch.Tensor, ws: int):
    B, H, W, C = x.shape
    ph, pw = (ws - H % ws) % ws, (ws - W % ws) % ws
    x = F.pad(x, (0, 0, 0, pw, 0, ph))
    Hp, Wp = H + ph, W + pw
    x = x.view(B, Hp // ws, ws, Wp // ws, ws, C).permute(0, 1, 3, 2, 4, 5).reshape(-1, ws, ws, C)
    return x, (Hp, Wp)


def window_unpartition(win: torch.Tensor, ws: int, pad_hw, hw):
    Hp, Wp = pad_hw
    H, W = hw
    B = win.shape[0] // (Hp * Wp // ws // ws)
    x = win.view(B, Hp // ws, Wp // ws, ws, ws, -1).permute(0, 1, 3, 2, 4, 5).reshape(B, Hp, Wp, -1)
    return x[:, :H, :W, :]


def sam_image_encoder(w: dict, cfg: dict, x: torch.Tensor) -> torch.Tensor:
    """x (B, 3, H, W) f32 in 0..1 (the reference feeds the encoder directly, without SAM's pixel normalisation,
    tfds_dense_descriptor.py:122-123) -> (B, out_chans, H/16, W/16)."""
    d, heads, ws = cfg["dim"], cfg["heads"], cfg["window"]
    x = F.conv2d(x, w["patch_embed.proj.weight"], w["patch_embed.proj.bias"], stride=cfg["patch"]).permute(0, 2, 3, 1)
    x = x + w["pos_embed"]
    H, W = x.shape[1:3]
    for i in range(cfg["depth"]):
        b = f"blocks.{i}."
        y = F.layer_norm(x, (d,), w[b + "norm1.weight"], w[b + "norm1.bias"], 1e-6)
        if i in cfg["global_attn"]:
            y = attention_relpos(y, w, b, heads)
        else:
            y, pad_hw = window_partition(y, ws)
            y = attention_relpos(y, w, b, heads)
            y = window_unpartition(y, ws, pad_hw, (H, W))
        x = x + y
        y = F.layer_norm(x, (d,), w[b + "norm2.weight"], w[b + "norm2.bias"], 1e-6)
        y = F.linear(F.gelu(F.linear(y, w[b + "mlp.lin1.weight"], w[b + "mlp.lin1.bias"])), w[b + "mlp.lin2.weight"], w[b + "mlp.lin2.bias"])
        x = x + y
    x = F.conv2d(x.permute(0, 3, 1, 2), w["neck.0.weight"])
    x = _layernorm2d(x, w["neck.1.weight"], w["neck.1.bias"])
    x = F.conv2d(x, w["neck.2.weight"], padding=1)
    return _layernorm2d(x, w["neck.3.weight"], w["neck.3.bias"])


def _layernorm2d(x, gamma, beta, eps=1e-6):
    u = x.mean(1, keepdim=True)
    s = (x - u).pow(2).mean(1, keepdim=True)
    return gamma[:, None, None] * ((x - u) / torch.sqrt(s + eps)) + beta[:, None, None]


def sam_dense_descriptor(w: dict, cfg: dict, x: torch.Tensor) -> torch.Tensor:
    """What get_dense_descriptor returns for 'medsam' (:123-126), batched: (B, H/16, W/16, out_chans)."""
    return sam_image_encoder(w, cfg, x).permute(0, 2, 3, 1)


# ----------------------------------------------------------------------------------------------- DINOv2 patch_embed mode
def init_dinov2_state_dict(dim: int = 384, patch: int = 14, seed: int = 1234) -> dict:
    g = torch.Generator().manual_seed(seed)
    wt = torch.empty(dim, 3, patch, patch)
    torch.nn.init.trunc_normal_(wt, std=0.02, a=-0.04, b=0.04, generator=g)
    bs = torch.empty(dim)
    torch.nn.init.trunc_normal_(bs, std=0.01, a=-0.02, b=0.02, generator=g)
    return {"patch_embed.proj.weight": wt, "patch_embed.proj.bias": bs}


def dinov2_patch_embed(w: dict, x: torch.Tensor) -> torch.Tensor:
    """x (B, 3, H, W) -> (B, H/14, W/14, dim): DINOv2 PatchEmbed.forward (conv, flatten, no norm) reshaped as
    get_dense_descriptor does (:128-133)."""
    return F.conv2d(x, w["patch_embed.proj.weight"], w["patch_embed.proj.bias"], stride=w["patch_embed.proj.weight"].shape[-1]).permute(0, 2, 3, 1)
