"""TEST INFRASTRUCTURE ONLY -- functional restatement of the reference's point-cloud classifier.

Oracle for SURVEY.md section 8(a) rows M1 (TransformerNoduleClassifier), M2 (FocalLoss).
The reference builds the model from ``torch.nn.TransformerEncoderLayer`` (post-norm,
norm_first=False, activation 'gelu', batch_first) -- third-party arithmetic, unpinned by the
reference; this image pins torch 2.11.0.  The restatement below spells the arithmetic out with
explicit matmul / softmax / layer_norm calls on a plain state-dict so the CUDA path can be
compared op by op; ``tests/test_oracle_vs_reference.py`` pins it against the UNMODIFIED
``models_archs.TransformerNoduleClassifier`` (eval mode, dropout inactive) and
``tests/golden/classifier_small.npz`` freezes the reference's own outputs and gradients.
Nothing in the product package imports this file.
"""
from __future__ import annotations

import math

import torch
import torch.nn.functional as F


def classifier_forward(sd: dict, x: torch.Tensor, num_heads: int, num_layers: int):
    """reference: src/models_archs.py:141-147 (+ MLPLayer.forward :193-199), dropout inactive.

    sd : state-dict with the reference's key names
    x  : (B, n, d)
    returns (logits (B, C), cls (B, d))
    """
    B, n, d = x.shape
    hd = d // num_heads
    cls = sd["cls_token"].expand(B, 1, d)                                         # :143
    t = torch.cat([cls, x], dim=1)                                                # :144
    t = F.layer_norm(t, (d,), sd["norm.weight"], sd["norm.bias"], eps=1e-5)       # :145
    N = n + 1
    for i in range(num_layers):                                                   # :146
        p = f"transformer_encoder.layers.{i}."
        qkv = F.linear(t, sd[p + "self_attn.in_proj_weight"], sd[p + "self_attn.in_proj_bias"])
        q, k, v = qkv.split(d, dim=-1)
        q = q.reshape(B, N, num_heads, hd).transpose(1, 2)
        k = k.reshape(B, N, num_heads, hd).transpose(1, 2)
        v = v.reshape(B, N, num_heads, hd).transpose(1, 2)
        a = torch.softmax((q @ k.transpose(-1, -2)) / math.sqrt(hd), dim=-1)
        o = (a @ v).transpose(1, 2).reshape(B, N, d)
        o = F.linear(o, sd[p + "self_attn.out_proj.weight"], sd[p + "self_attn.out_proj.bias"])
        t = F.layer_norm(t + o, (d,), sd[p + "norm1.weight"], sd[p + "norm1.bias"], eps=1e-5)
        y = F.gelu(F.linear(t, sd[p + "linear1.weight"], sd[p + "linear1.bias"]))
        y = F.linear(y, sd[p + "linear2.weight"], sd[p + "linear2.bias"])
        t = F.layer_norm(t + y, (d,), sd[p + "norm2.weight"], sd[p + "norm2.bias"], eps=1e-5)
    cls_out = t[:, 0, :]                                                          # :147
    y = F.gelu(F.linear(cls_out, sd["classifier.dense1.weight"], sd["classifier.dense1.bias"]))
    logits = F.linear(y, sd["classifier.dense2.weight"], sd["classifier.dense2.bias"])
    return logits, cls_out


def focal_loss(logits: torch.Tensor, targets_onehot: torch.Tensor, gamma=2.0, alpha=None):
    """reference: src/train_models.py:390-405 -- sum over the batch of
    -alpha[c] * (1 - p_c)^gamma * log p_c at the true class c = argmax(one-hot)."""
    if logits.dim() == 1:
        logits = logits.unsqueeze(0)
        targets_onehot = targets_onehot.unsqueeze(0)
    c = torch.argmax(targets_onehot, dim=1)
    logpt = F.log_softmax(logits, dim=1)
    pt = torch.exp(logpt)
    mod = (1 - pt) ** gamma * logpt
    picked = mod.gather(1, c[:, None])[:, 0]
    wgt = torch.ones_like(picked) if alpha is None else alpha.to(picked.dtype)[c]
    return -(wgt * picked).sum()


def init_state_dict(input_dim=256, dim_feedforward=1024, num_classes=2, num_layers=2,
                    seed=1234, dtype=torch.float32) -> dict:
    """Seeded random state-dict with the reference's key names and shapes (section 3.3).
    Uses its own generator so it is reproducible without the reference being importable."""
    g = torch.Generator().manual_seed(seed)
    d, ff = input_dim, dim_feedforward

    def rn(*shape, std):
        return (torch.randn(*shape, generator=g) * std).to(dtype)

    sd = {"cls_token": rn(1, 1, d, std=1.0),
          "norm.weight": 1 + rn(d, std=0.05), "norm.bias": rn(d, std=0.02)}
    for i in range(num_layers):
        p = f"transformer_encoder.layers.{i}."
        sd[p + "self_attn.in_proj_weight"] = rn(3 * d, d, std=1 / math.sqrt(d))
        sd[p + "self_attn.in_proj_bias"] = rn(3 * d, std=0.02)
        sd[p + "self_attn.out_proj.weight"] = rn(d, d, std=1 / math.sqrt(d))
        sd[p + "self_attn.out_proj.bias"] = rn(d, std=0.02)
        sd[p + "linear1.weight"] = rn(ff, d, std=1 / math.sqrt(d))
        sd[p + "linear1.bias"] = rn(ff, std=0.02)
        sd[p + "linear2.weight"] = rn(d, ff, std=1 / math.sqrt(ff))
        sd[p + "linear2.bias"] = rn(d, std=0.02)
        for nm in ("norm1", "norm2"):
            sd[p + nm + ".weight"] = 1 + rn(d, std=0.05)
            sd[p + nm + ".bias"] = rn(d, std=0.02)
    sd["classifier.dense1.weight"] = rn(2 * d, d, std=1 / math.sqrt(d))
    sd["classifier.dense1.bias"] = rn(2 * d, std=0.02)
    sd["classifier.dense2.weight"] = rn(num_classes, 2 * d, std=1 / math.sqrt(2 * d))
    sd["classifier.dense2.bias"] = rn(num_classes, std=0.02)
    return sd


def _encoder(sd, prefix_cls, prefix_norm, prefix_enc, x, num_heads, num_layers):
    """cat(cls, x) -> LayerNorm -> post-norm encoder layers (the body of classifier_forward with the bimodal key prefixes)."""
    B, n, d = x.shape
    hd = d // num_heads
    t = torch.cat([sd[prefix_cls].expand(B, 1, d), x], dim=1)
    t = F.layer_norm(t, (d,), sd[prefix_norm + ".weight"], sd[prefix_norm + ".bias"], eps=1e-5)
    N = n + 1
    for i in range(num_layers):
        p = f"{prefix_enc}.layers.{i}."
        qkv = F.linear(t, sd[p + "self_attn.in_proj_weight"], sd[p + "self_attn.in_proj_bias"])
        q, k, v = qkv.split(d, dim=-1)
        q = q.reshape(B, N, num_heads, hd).transpose(1, 2)
        k = k.reshape(B, N, num_heads, hd).transpose(1, 2)
        v = v.reshape(B, N, num_heads, hd).transpose(1, 2)
        a = torch.softmax((q @ k.transpose(-1, -2)) / math.sqrt(hd), dim=-1)
        o = (a @ v).transpose(1, 2).reshape(B, N, d)
        o = F.linear(o, sd[p + "self_attn.out_proj.weight"], sd[p + "self_attn.out_proj.bias"])
        t = F.layer_norm(t + o, (d,), sd[p + "norm1.weight"], sd[p + "norm1.bias"], eps=1e-5)
        y = F.gelu(F.linear(t, sd[p + "linear1.weight"], sd[p + "linear1.bias"]))
        y = F.linear(y, sd[p + "linear2.weight"], sd[p + "linear2.bias"])
        t = F.layer_norm(t + y, (d,), sd[p + "norm2.weight"], sd[p + "norm2.bias"], eps=1e-5)
    return t


def _mlp(sd, prefix, x):
    return F.linear(F.gelu(F.linear(x, sd[prefix + ".dense1.weight"], sd[prefix + ".dense1.bias"])),
                    sd[prefix + ".dense2.weight"], sd[prefix + ".dense2.bias"])


def _cross(sd, prefix, q_in, kv_in, num_heads):
    """nn.MultiheadAttention(batch_first) restated: packed in_proj, per-head softmax(q k^T / sqrt(hd)) v, out_proj."""
    d = q_in.shape[-1]
    hd = d // num_heads
    w, b = sd[prefix + ".multihead_attn.in_proj_weight"], sd[prefix + ".multihead_attn.in_proj_bias"]
    q = F.linear(q_in, w[:d], b[:d])
    k = F.linear(kv_in, w[d:2 * d], b[d:2 * d])
    v = F.linear(kv_in, w[2 * d:], b[2 * d:])
    B, Nq, Nk = q.shape[0], q.shape[1], k.shape[1]
    q = q.reshape(B, Nq, num_heads, hd).transpose(1, 2)
    k = k.reshape(B, Nk, num_heads, hd).transpose(1, 2)
    v = v.reshape(B, Nk, num_heads, hd).transpose(1, 2)
    a = torch.softmax((q @ k.transpose(-1, -2)) / math.sqrt(hd), dim=-1)
    o = (a @ v).transpose(1, 2).reshape(B, Nq, d)
    return F.linear(o, sd[prefix + ".multihead_attn.out_proj.weight"], sd[prefix + ".multihead_attn.out_proj.bias"])


def bimodal_forward(sd: dict, x_ct, x_pet, heads_ct, heads_pet, layers_ct, layers_pet):
    """reference: src/models_archs.py:76-124 (TransformerNoduleBimodalClassifier.forward), dropout inactive.
    Returns (logits_petct, petct_cls_token, logits_ct, logits_pet)."""
    use_ct, use_pet = x_ct is not None, x_pet is not None
    assert use_ct or use_pet
    if use_ct:
        t_ct = _encoder(sd, "cls_token_ct", "norm_ct", "transformer_encoder_ct", x_ct, heads_ct, layers_ct)
        ct_cls = t_ct[:, 0, :]
    if use_pet:
        t_pet = _encoder(sd, "cls_token_pet", "norm_pet", "transformer_encoder_pet", x_pet, heads_pet, layers_pet)
        pet_cls = t_pet[:, 0, :]
    if use_ct and use_pet:
        ct_cls = _cross(sd, "cross_attention_ct", t_ct, t_pet, heads_ct)[:, 0, :]            # :100-103
        pet_cls = _cross(sd, "cross_attention_pet", t_pet, t_ct, heads_ct)[:, 0, :]          # built with num_heads_ct (:71)
        logits_ct = _mlp(sd, "classifier_ct", ct_cls)
        logits_pet = _mlp(sd, "classifier_pet", pet_cls)
        z = _mlp(sd, "projection_petct", torch.cat([ct_cls, pet_cls], dim=1))
        return _mlp(sd, "classifier_petct", z), z, logits_ct, logits_pet
    if use_ct:
        lg = _mlp(sd, "classifier_ct", ct_cls)
        return lg, ct_cls, lg, lg
    lg = _mlp(sd, "classifier_pet", pet_cls)
    return lg, pet_cls, lg, lg
