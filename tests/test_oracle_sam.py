"""oracle/sam_fp32.py (restatement of segment_anything's ImageEncoderViT, the backbone the reference loads for 'medsam',
src/tfds_dense_descriptor.py:104,123) against an independent implementation of the same architecture:
transformers' SamVisionEncoder with the same weights.  CPU only."""
import pytest
import torch

from oracle import sam_fp32


def _to_hf(sd: dict) -> dict:
    """segment_anything key names -> transformers SamVisionEncoder key names."""
    out = {}
    for k, v in sd.items():
        k2 = (k.replace("patch_embed.proj.", "patch_embed.projection.").replace("blocks.", "layers.")
               .replace(".norm1.", ".layer_norm1.").replace(".norm2.", ".layer_norm2.")
               .replace("neck.0.", "neck.conv1.").replace("neck.1.", "neck.layer_norm1.")
               .replace("neck.2.", "neck.conv2.").replace("neck.3.", "neck.layer_norm2."))
        out[k2] = v
    return out


@pytest.mark.parametrize("img", [256, 224])
def test_sam_oracle_matches_hf_encoder(img):
    from transformers import SamVisionConfig
    from transformers.models.sam.modeling_sam import SamVisionEncoder
    cfg = sam_fp32.SAM_CONFIGS["sam_tiny"]
    sd = sam_fp32.init_sam_state_dict(cfg, (img, img), seed=5)
    hf_cfg = SamVisionConfig(hidden_size=cfg["dim"], output_channels=cfg["out_chans"], num_hidden_layers=cfg["depth"],
                             num_attention_heads=cfg["heads"], image_size=img, patch_size=16, window_size=cfg["window"],
                             global_attn_indexes=list(cfg["global_attn"]), mlp_dim=4 * cfg["dim"])
    enc = SamVisionEncoder(hf_cfg).eval()
    missing, unexpected = enc.load_state_dict(_to_hf(sd), strict=True)
    assert not missing and not unexpected
    x = torch.rand(2, 3, img, img, generator=torch.Generator().manual_seed(1))
    with torch.no_grad():
        want = enc(x).last_hidden_state
        got = sam_fp32.sam_image_encoder(sd, cfg, x)
    assert got.shape == want.shape == (2, cfg["out_chans"], img // 16, img // 16)
    assert (got - want).abs().max() < 2e-4, float((got - want).abs().max())


def test_window_partition_round_trip_and_padding():
    x = torch.arange(2 * 16 * 16 * 3, dtype=torch.float32).reshape(2, 16, 16, 3) + 1
    win, pad_hw = sam_fp32.window_partition(x, 14)
    assert win.shape == (2 * 4, 14, 14, 3) and pad_hw == (28, 28)
    assert (win[1, :, 2:] == 0).all() and (win[2, 2:] == 0).all()     # right / bottom padding is zero
    assert torch.equal(sam_fp32.window_unpartition(win, 14, pad_hw, (16, 16)), x)


def test_dinov2_patch_embed_shape():
    w = sam_fp32.init_dinov2_state_dict()
    out = sam_fp32.dinov2_patch_embed(w, torch.rand(1, 3, 56, 56))
    assert out.shape == (1, 4, 4, 384)


def test_sam_oracle_against_committed_golden(golden_dir):
    """tests/golden/sam_tiny_hf.npz: SamVisionEncoder output frozen by tests/golden/make_golden_sam.py."""
    import importlib.util
    import os
    import numpy as np
    spec = importlib.util.spec_from_file_location("make_golden_sam", os.path.join(golden_dir, "make_golden_sam.py"))
    mk = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mk)
    g = np.load(os.path.join(golden_dir, "sam_tiny_hf.npz"))
    cfg = sam_fp32.SAM_CONFIGS["sam_tiny"]
    sd = sam_fp32.init_sam_state_dict(cfg, (mk.IMG, mk.IMG), seed=mk.SEED_W)
    with torch.no_grad():
        got = sam_fp32.sam_dense_descriptor(sd, cfg, mk.golden_input())[0].numpy()
    assert got.shape == g["descriptors"].shape == (16, 16, 64)
    assert np.abs(got - g["descriptors"]).max() < 2e-4
