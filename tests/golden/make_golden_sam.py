"""Freeze a golden vector for the MedSAM image-encoder path.

The reference's backbone is third-party code it neither vendors nor pins (segment_anything, src/tfds_dense_descriptor.py:104),
so there is nothing of the reference to run.  The golden output is produced by an INDEPENDENT implementation of the same
architecture -- transformers' SamVisionEncoder (transformers 5.5, torch 2.11, CPU fp32) -- fed the seeded weights of
oracle.sam_fp32.init_sam_state_dict renamed to its key names:

    python tests/golden/make_golden_sam.py        # writes tests/golden/sam_tiny_hf.npz (65 KB)

Inputs are regenerated from the seeds in the tests; only the expected output is stored.
"""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import sam_fp32  # noqa: E402

IMG, SEED_W, SEED_X = 256, 21, 22


def golden_input():
    return torch.rand(1, 3, IMG, IMG, generator=torch.Generator().manual_seed(SEED_X))


def to_hf(sd):
    out = {}
    for k, v in sd.items():
        k2 = (k.replace("patch_embed.proj.", "patch_embed.projection.").replace("blocks.", "layers.")
               .replace(".norm1.", ".layer_norm1.").replace(".norm2.", ".layer_norm2.")
               .replace("neck.0.", "neck.conv1.").replace("neck.1.", "neck.layer_norm1.")
               .replace("neck.2.", "neck.conv2.").replace("neck.3.", "neck.layer_norm2."))
        out[k2] = v
    return out


if __name__ == "__main__":
    from transformers import SamVisionConfig
    from transformers.models.sam.modeling_sam import SamVisionEncoder
    cfg = sam_fp32.SAM_CONFIGS["sam_tiny"]
    sd = sam_fp32.init_sam_state_dict(cfg, (IMG, IMG), seed=SEED_W)
    enc = SamVisionEncoder(SamVisionConfig(hidden_size=cfg["dim"], output_channels=cfg["out_chans"], num_hidden_layers=cfg["depth"],
                                           num_attention_heads=cfg["heads"], image_size=IMG, patch_size=16, window_size=cfg["window"],
                                           global_attn_indexes=list(cfg["global_attn"]), mlp_dim=4 * cfg["dim"])).eval()
    enc.load_state_dict(to_hf(sd), strict=True)
    with torch.no_grad():
        out = enc(golden_input()).last_hidden_state[0].permute(1, 2, 0).contiguous().numpy()     # (16, 16, 64) HWC as get_dense_descriptor returns
    path = os.path.join(os.path.dirname(os.path.abspath(__file__)), "sam_tiny_hf.npz")
    np.savez_compressed(path, descriptors=out.astype(np.float32), img=IMG, seed_w=SEED_W, seed_x=SEED_X)
    print(path, out.shape, os.path.getsize(path))
