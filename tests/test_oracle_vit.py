"""oracle/vit_fp32.py against transformers.ViTModel with copied weights (CPU)."""
import pytest
import torch

from oracle import vit_fp32


@pytest.mark.parametrize("patch,H,W", [(16, 64, 64), (14, 56, 84), (16, 48, 80)])
def test_vit_oracle_matches_hf(patch, H, W):
    """patch 16 square (C1 / C2 geometry), patch 14 and non-square grids (C4's ViT-L/14 geometry; crops resized to other inputs)."""
    tr = pytest.importorskip("transformers")
    cfg = dict(dim=128, depth=2, heads=2, patch=patch)
    w = vit_fp32.init_weights(cfg, (H, W), seed=3)
    hf_cfg = tr.ViTConfig(hidden_size=128, num_hidden_layers=2, num_attention_heads=2, intermediate_size=512,
                          image_size=(H, W), patch_size=patch, hidden_act="gelu", layer_norm_eps=1e-6, qkv_bias=True,
                          hidden_dropout_prob=0.0, attention_probs_dropout_prob=0.0)
    m = tr.ViTModel(hf_cfg, add_pooling_layer=False).eval()
    sd = m.state_dict()
    d = 128
    sd["embeddings.cls_token"] = w["cls_token"]
    sd["embeddings.position_embeddings"] = w["pos_embed"]
    sd["embeddings.patch_embeddings.projection.weight"] = w["patch_embed.weight"]
    sd["embeddings.patch_embeddings.projection.bias"] = w["patch_embed.bias"]
    for i in range(2):
        b, hb = f"blocks.{i}.", f"encoder.layer.{i}."
        qkv_w, qkv_b = w[b + "attn.qkv.weight"], w[b + "attn.qkv.bias"]
        for j, nm in enumerate(("query", "key", "value")):
            sd[hb + f"attention.attention.{nm}.weight"] = qkv_w[j * d:(j + 1) * d]
            sd[hb + f"attention.attention.{nm}.bias"] = qkv_b[j * d:(j + 1) * d]
        sd[hb + "attention.output.dense.weight"] = w[b + "attn.proj.weight"]
        sd[hb + "attention.output.dense.bias"] = w[b + "attn.proj.bias"]
        sd[hb + "layernorm_before.weight"], sd[hb + "layernorm_before.bias"] = w[b + "norm1.weight"], w[b + "norm1.bias"]
        sd[hb + "layernorm_after.weight"], sd[hb + "layernorm_after.bias"] = w[b + "norm2.weight"], w[b + "norm2.bias"]
        sd[hb + "intermediate.dense.weight"], sd[hb + "intermediate.dense.bias"] = w[b + "mlp.fc1.weight"], w[b + "mlp.fc1.bias"]
        sd[hb + "output.dense.weight"], sd[hb + "output.dense.bias"] = w[b + "mlp.fc2.weight"], w[b + "mlp.fc2.bias"]
    sd["layernorm.weight"], sd["layernorm.bias"] = w["norm.weight"], w["norm.bias"]
    m.load_state_dict(sd)
    x = torch.randn(2, 3, H, W)
    with torch.no_grad():
        want = m(pixel_values=x).last_hidden_state
        dense, tok = vit_fp32.vit_forward(w, cfg, x, return_tokens=True)
    assert torch.allclose(tok, want, atol=2e-5, rtol=1e-4)
    gh, gw = H // patch, W // patch
    assert dense.shape == (2, gh, gw, 128) and torch.equal(dense.reshape(2, gh * gw, 128), tok[:, 1:])


def test_gray2rgb_patch_embedding_equals_channel_summed_weights():
    """The identity behind vdr_patch_embed_gemm_gray (include/vdr.h): a gray slice fed to the three input channels (gray2rgb,
    reference tfds_dense_descriptor.py:41) through Conv2d(3, d, p, p) == the same slice through the weights summed over the input
    channels -- in f64 exactly, in f32 to rounding; and the oracle's whole forward agrees when its patch weights are replaced by
    (W_r + W_g + W_b) / 3 in every channel."""
    import torch.nn.functional as F
    cfg = vit_fp32.VIT_CONFIGS["vit_t16"] if "vit_t16" in vit_fp32.VIT_CONFIGS else next(iter(vit_fp32.VIT_CONFIGS.values()))
    hw = (64, 64)
    w = vit_fp32.init_weights(cfg, hw, seed=5)
    g = torch.rand(2, 1, *hw, generator=torch.Generator().manual_seed(1))
    x3 = g.expand(-1, 3, -1, -1).contiguous()
    W, b, p = w["patch_embed.weight"], w["patch_embed.bias"], cfg["patch"]
    a = F.conv2d(x3.double(), W.double(), b.double(), stride=p)
    c = F.conv2d(g.double(), W.double().sum(dim=1, keepdim=True), b.double(), stride=p)
    assert torch.allclose(a, c, rtol=0, atol=1e-12)
    a32 = F.conv2d(x3, W, b, stride=p)
    c32 = F.conv2d(g, W.sum(dim=1, keepdim=True), b, stride=p)
    assert (a32 - c32).abs().max() < 1e-5
    w2 = dict(w)
    w2["patch_embed.weight"] = (W.sum(dim=1, keepdim=True) / 3).expand(-1, 3, -1, -1).contiguous()
    with torch.no_grad():
        d0 = vit_fp32.vit_forward(w, cfg, x3)
        d1 = vit_fp32.vit_forward(w2, cfg, x3)
    assert (d0 - d1).abs().max() < 1e-4
