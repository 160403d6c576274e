"""-m gpu: the MedSAM / SAM image-encoder path and the reference's 'dinov2' patch-embedding mode
against the oracle (oracle/sam_fp32.py, pinned to transformers' SamVisionEncoder) and the committed golden vector.

Tolerances: integer / byte work (window partition, 3x3 im2col) bit-exact; bf16 activations with fp32 accumulation against
the fp32 oracle: |error| <= ABS_TOL on the LayerNorm-ed descriptors (O(1) values), rms relative error <= REL_TOL,
per-descriptor cosine >= 0.999 (BASELINE.json north_star)."""
import importlib.util
import math
import os

import numpy as np
import pytest
import torch
import torch.nn.functional as F

pytestmark = pytest.mark.gpu

ABS_TOL, REL_TOL, COS = 0.15, 0.02, 0.999        # fallback; the per-config stated bounds are in tests/parity.py


def _check_descriptors(got, want, key):
    import parity
    return parity.check(key, got, want, dict(abs=ABS_TOL, rel=REL_TOL, cos=COS))


@pytest.mark.parametrize("B,H,W,ws,d", [(2, 16, 16, 14, 128), (1, 14, 14, 14, 64), (3, 9, 20, 7, 72), (1, 64, 64, 14, 768)])
def test_window_rows_bit_exact(cuda, B, H, W, ws, d):
    from oracle import sam_fp32
    from vit_deep_radiomics_b200 import ops
    x = torch.randn(B, H, W, d, generator=torch.Generator().manual_seed(1)).bfloat16()
    want, pad_hw = sam_fp32.window_partition(x.float(), ws)
    got = ops.window_rows(x.reshape(-1, d).to(cuda), B, H, W, ws, True)
    assert torch.equal(got.cpu().float().reshape(want.shape), want)
    back = ops.window_rows(got, B, H, W, ws, False)
    assert torch.equal(back.cpu().reshape(B, H, W, d), x)
    with pytest.raises(ValueError):
        ops.window_rows(x.reshape(-1, d)[1:].to(cuda), B, H, W, ws, True)


def _attn_reference(qkv, BW, Sh, Sw, heads, rel_h, rel_w):
    """fp32 attention with the decomposed bias on the bf16-rounded operands (segment_anything Attention.forward)."""
    N = Sh * Sw
    q, k, v = qkv.float().reshape(BW, N, 3, heads, 64).permute(2, 0, 3, 1, 4).reshape(3, BW * heads, N, 64).unbind(0)
    ih = torch.arange(Sh)[:, None] - torch.arange(Sh)[None, :] + Sh - 1
    iw = torch.arange(Sw)[:, None] - torch.arange(Sw)[None, :] + Sw - 1
    rq = q.reshape(BW * heads, Sh, Sw, 64)
    bh = torch.einsum("bhwc,hkc->bhwk", rq, rel_h[ih])
    bw = torch.einsum("bhwc,wkc->bhwk", rq, rel_w[iw])
    a = (q * 0.125) @ k.transpose(-2, -1)
    a = (a.view(BW * heads, Sh, Sw, Sh, Sw) + bh[..., :, None] + bw[..., None, :]).view(BW * heads, N, N).softmax(-1)
    out = (a @ v).view(BW, heads, N, 64).permute(0, 2, 1, 3).reshape(BW * N, heads * 64)
    return out, torch.cat([bh, bw], dim=-1).reshape(BW * heads * N, Sh + Sw)


@pytest.mark.parametrize("BW,Sh,Sw,heads", [(3, 14, 14, 2), (2, 16, 16, 2), (1, 8, 12, 3), (1, 5, 3, 1), (2, 7, 1, 1), (1, 64, 64, 2), (2, 3, 64, 1), (50, 14, 14, 12), (1, 40, 72, 1)])
def test_attn_relpos_vs_fp32(cuda, BW, Sh, Sw, heads):
    """Windowed (14x14 = 196 tokens, a ragged last key tile), global (64x64 = 4096) and non-square extents."""
    from vit_deep_radiomics_b200 import _C, ops
    g = torch.Generator().manual_seed(BW * 1000 + Sh * 10 + Sw)
    N, d = Sh * Sw, heads * 64
    qkv = (torch.randn(BW * N, 3 * d, generator=g) * 1.2).bfloat16()
    rel_h = torch.randn(2 * Sh - 1, 64, generator=g) * 0.1
    rel_w = torch.randn(2 * Sw - 1, 64, generator=g) * 0.1
    want, rel_want = _attn_reference(qkv, BW, Sh, Sw, heads, rel_h, rel_w)
    hi, lo = ops.relpos_split(rel_h.to(cuda), rel_w.to(cuda))
    assert (hi.float() + lo.float() - torch.cat([rel_h, rel_w]).to(cuda)).abs().max() < 2e-5
    rel_got = ops.relpos_tables(qkv.to(cuda), BW, Sh, Sw, heads, hi, lo).cpu()
    assert torch.allclose(rel_got, rel_want, atol=3e-4, rtol=1e-4), float((rel_got - rel_want).abs().max())
    n0 = _C.launch_count()
    got = ops.attn_relpos(qkv.to(cuda), BW, Sh, Sw, heads, hi, lo, kernel="mma")
    assert _C.launch_count() - n0 == 1            # bias terms are built inside the attention kernel
    got = got.cpu().float()
    assert torch.isfinite(got).all()
    err = (got - want).abs().max()
    cos = F.cosine_similarity(got.reshape(-1, 64), want.reshape(-1, 64), dim=1).min()
    assert err < 2e-2 and cos > 0.9995, (float(err), float(cos))


@pytest.mark.parametrize("BW,Sh,heads", [(1, 64, 2), (2, 64, 12), (3, 4, 1), (1, 8, 3)])
def test_flash_attn_relpos_tcgen05_vs_fp32(cuda, BW, Sh, heads):
    """The tcgen05 flash kernel with the bias (token grids of Sh x 64): against fp32 and against the mma.sync kernel."""
    from vit_deep_radiomics_b200 import _C, ops
    Sw = 64
    g = torch.Generator().manual_seed(77 + BW + Sh)
    N, d = Sh * Sw, heads * 64
    qkv = (torch.randn(BW * N, 3 * d, generator=g) * 1.2).bfloat16()
    rel_h = torch.randn(2 * Sh - 1, 64, generator=g) * 0.1
    rel_w = torch.randn(2 * Sw - 1, 64, generator=g) * 0.1
    want, _ = _attn_reference(qkv, BW, Sh, Sw, heads, rel_h, rel_w)
    hi, lo = ops.relpos_split(rel_h.to(cuda), rel_w.to(cuda))
    got = ops.attn_relpos(qkv.to(cuda), BW, Sh, Sw, heads, hi, lo, kernel="tcgen05").cpu().float()
    ref2 = ops.attn_relpos(qkv.to(cuda), BW, Sh, Sw, heads, hi, lo, kernel="mma").cpu().float()
    assert torch.isfinite(got).all()
    err = (got - want).abs().max()
    cos = F.cosine_similarity(got.reshape(-1, 64), want.reshape(-1, 64), dim=1).min()
    assert err < 2e-2 and cos > 0.9995, (float(err), float(cos))
    assert (got - ref2).abs().max() < 2e-2
    # the same kernel computing its own bias terms (no table, one launch): what the encoder runs
    n0 = _C.launch_count()
    fused = ops.attn_relpos(qkv.to(cuda), BW, Sh, Sw, heads, hi, lo, kernel="fused").cpu().float()
    assert _C.launch_count() - n0 == 1
    assert torch.isfinite(fused).all()
    err = (fused - want).abs().max()
    cos = F.cosine_similarity(fused.reshape(-1, 64), want.reshape(-1, 64), dim=1).min()
    assert err < 2e-2 and cos > 0.9995, (float(err), float(cos))
    assert (fused - got).abs().max() < 2e-2            # one bf16 ulp of an output of magnitude 2..4: rel_w is fp16 in both, added by the MMA here
    with pytest.raises(ValueError):      # Sh = 2: not a multiple of 4 -> the tcgen05 variant refuses, it never falls back silently
        hi2, lo2 = ops.relpos_split(torch.zeros(3, 64, device=cuda), torch.zeros(127, 64, device=cuda))
        ops.attn_relpos(qkv.to(cuda)[:128], 1, 2, Sw, heads, hi2, lo2, kernel="tcgen05")


@pytest.mark.parametrize("B,gh,gw,ws,heads", [(2, 16, 16, 14, 2), (1, 64, 64, 14, 12), (3, 9, 20, 7, 1), (1, 14, 14, 14, 2)])
def test_attn_relpos_windows_in_place_equals_partitioned_path(cuda, B, gh, gw, ws, heads):
    """Windows read in place (pad tokens = the qkv bias) against window_partition -> attention -> window_unpartition: bit for bit
    on the mma.sync kernel (ws != 14), to bf16 rounding on the tcgen05 kernel (ws == 14)."""
    from vit_deep_radiomics_b200 import ops
    g = torch.Generator().manual_seed(B + gh + gw)
    d = heads * 64
    bias = (torch.randn(3 * d, generator=g) * 0.3).to(cuda)
    qkv = (torch.randn(B * gh * gw, 3 * d, generator=g) * 1.1).bfloat16().to(cuda)
    hi, lo = ops.relpos_split((torch.randn(2 * ws - 1, 64, generator=g) * 0.1).to(cuda), (torch.randn(2 * ws - 1, 64, generator=g) * 0.1).to(cuda))
    got = ops.attn_relpos_windows(qkv, bias, B, gh, gw, ws, heads, hi, lo)
    # partitioned path: pad rows of the windowed qkv matrix hold the bias (what the qkv GEMM writes for a zero input row)
    qw = ops.window_rows(qkv, B, gh, gw, ws, True)
    ones = ops.window_rows(torch.ones(B * gh * gw, 8, dtype=torch.bfloat16, device=cuda), B, gh, gw, ws, True)[:, 0]
    qw[ones == 0] = bias.bfloat16()
    nwin = (-(-gh // ws)) * (-(-gw // ws))
    ow = ops.attn_relpos(qw, B * nwin, ws, ws, heads, hi, lo, kernel="mma")
    want = ops.window_rows(ow, B, gh, gw, ws, False)
    if ws == 14:
        # SAM's own window size runs on the tcgen05 kernel (sam_window_tc.cu): same operands, different accumulation order and
        # bias folded into the score MMA as bf16 hi + lo parts -> agreement to the rounding of the bf16 outputs, and against fp32
        assert torch.isfinite(got.float()).all()
        err = (got.float() - want.float()).abs().max()
        cos = F.cosine_similarity(got.float().reshape(-1, 64), want.float().reshape(-1, 64), dim=1).min()
        assert err < 1.6e-2 and cos > 0.9997, (float(err), float(cos))
        rel_h, rel_w = (hi.float() + lo.float())[:2 * ws - 1].cpu(), (hi.float() + lo.float())[2 * ws - 1:].cpu()
        ref, _ = _attn_reference(qw.cpu(), B * nwin, ws, ws, heads, rel_h, rel_w)
        ref = ops.window_rows(ref.to(cuda).bfloat16().reshape(-1, d).contiguous(), B, gh, gw, ws, False).float()
        err32 = (got.float() - ref).abs().max()
        assert err32 < 2e-2, float(err32)
    else:
        assert torch.equal(got, want)


def test_attn_relpos_zero_bias_matches_flash_attention(cuda):
    """With zero rel-pos tables the kernel computes plain softmax attention: same result as the tcgen05 flash kernel."""
    from vit_deep_radiomics_b200 import ops
    g = torch.Generator().manual_seed(9)
    BW, S, heads = 2, 16, 2
    qkv = torch.randn(BW * S * S, 3 * heads * 64, generator=g).bfloat16().to(cuda)
    z = torch.zeros(2 * (2 * S - 1), 64, device=cuda, dtype=torch.bfloat16)
    a = ops.attn_relpos(qkv, BW, S, S, heads, z, z).float()
    b = ops.flash_attn(qkv, BW, S * S, heads).float()
    assert (a - b).abs().max() < 1.6e-2


@pytest.mark.parametrize("B,H,W,C", [(2, 16, 16, 64), (1, 5, 7, 8), (1, 64, 64, 256)])
def test_im2col3x3_bit_exact(cuda, B, H, W, C):
    from vit_deep_radiomics_b200 import ops
    x = torch.randn(B, H, W, C, generator=torch.Generator().manual_seed(3)).bfloat16()
    got = ops.im2col3x3_tokens(x.reshape(-1, C).to(cuda), B, H, W).cpu()
    # F.unfold orders (c, ky, kx); the kernel writes (ky, kx, c)
    u = F.unfold(x.float().permute(0, 3, 1, 2), 3, padding=1).reshape(B, C, 9, H * W).permute(0, 3, 2, 1).reshape(B * H * W, 9 * C)
    assert torch.equal(got.float(), u)


def test_neck_conv3x3_as_gemm(cuda):
    from vit_deep_radiomics_b200 import ops
    g = torch.Generator().manual_seed(4)
    B, H, W, C = 2, 16, 16, 64
    x = torch.randn(B, H, W, C, generator=g).bfloat16()
    wt = (torch.randn(C, C, 3, 3, generator=g) * 0.05).bfloat16()
    A = ops.im2col3x3_tokens(x.reshape(-1, C).to(cuda), B, H, W)
    got = ops.gemm(A, wt.permute(0, 2, 3, 1).reshape(C, 9 * C).contiguous().to(cuda), None, out_dtype=torch.float32).cpu()
    want = F.conv2d(x.float().permute(0, 3, 1, 2), wt.float(), padding=1).permute(0, 2, 3, 1).reshape(-1, C)
    assert torch.allclose(got, want, atol=2e-3, rtol=2e-3), float((got - want).abs().max())


@pytest.mark.parametrize("hw,B", [((256, 256), 2), ((224, 224), 1), ((192, 320), 1)])
def test_sam_encoder_vs_oracle(cuda, hw, B):
    """sam_tiny (4 blocks, windowed + global attention, neck): padded windows (16 -> 28), exact windows (14), non-square."""
    from oracle import sam_fp32
    from vit_deep_radiomics_b200 import _C, tfds_dense_descriptor as tdd
    model = tdd.load_model("sam_tiny", img_hw=hw, device=cuda, seed=13)
    x = torch.rand(B, 3, *hw, generator=torch.Generator().manual_seed(2))
    n0 = _C.launch_count()
    got = model.dense_descriptors(x.to(cuda)).cpu().numpy()
    assert _C.launch_count() > n0
    with torch.no_grad():
        want = sam_fp32.sam_dense_descriptor(model.state_dict_f32, model.cfg, x).numpy()
    assert got.shape == want.shape == (B, hw[0] // 16, hw[1] // 16, 64)
    _check_descriptors(got, want, f"descriptors sam_tiny@{hw[0]}x{hw[1]}")


def test_sam_encoder_layernorm_kernel_path_and_mma_global_attention(cuda):
    """The A/B switches: LayerNorm kernels instead of the folded GEMM epilogues, mma.sync instead of tcgen05 for a 64-wide grid."""
    from oracle import sam_fp32
    from vit_deep_radiomics_b200 import sam_encoder
    hw = (128, 1024)                                           # 8 x 64 tokens: the global blocks qualify for the tcgen05 kernel
    x = torch.rand(1, 3, *hw, generator=torch.Generator().manual_seed(12))
    outs = []
    for fold, kern, in_place in [(True, "auto", True), (False, "mma", True), (False, "auto", False)]:
        model = sam_encoder.SamImageEncoder("sam_tiny", img_hw=hw, device=cuda, seed=29)
        model.fold_layernorm, model.global_attn_kernel, model.windows_in_place = fold, kern, in_place
        model.prepare()
        outs.append(model.dense_descriptors(x.to(cuda)).cpu().numpy())
    with torch.no_grad():
        want = sam_fp32.sam_dense_descriptor(model.state_dict_f32, model.cfg, x).numpy()
    for got in outs:
        _check_descriptors(got, want, "descriptors sam_tiny@128x1024 (kernel variants)")


def test_sam_encoder_against_committed_golden(cuda, golden_dir):
    """tests/golden/sam_tiny_hf.npz = transformers' SamVisionEncoder on the same seeded weights / input."""
    from vit_deep_radiomics_b200 import sam_encoder
    spec = importlib.util.spec_from_file_location("make_golden_sam", os.path.join(golden_dir, "make_golden_sam.py"))
    mk = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mk)
    g = np.load(os.path.join(golden_dir, "sam_tiny_hf.npz"))
    cfg = sam_encoder.SAM_CONFIGS["sam_tiny"]
    sd = sam_encoder.init_sam_state_dict(cfg, (mk.IMG, mk.IMG), seed=mk.SEED_W)
    full = {"image_encoder." + k: v for k, v in sd.items()}                      # as a full SAM checkpoint stores it
    model = sam_encoder.SamImageEncoder("sam_tiny", img_hw=(mk.IMG, mk.IMG), state_dict=full, device=cuda)
    got = model.dense_descriptors(mk.golden_input().to(cuda))[0].cpu().numpy()
    _check_descriptors(got, g["descriptors"], "descriptors sam_tiny vs HF golden")


def test_medsam_full_size_slice(cuda):
    """The reference's actual configuration: ViT-B, 1024 x 1024 gray slice -> (64, 64, 256) (tfds_dense_descriptor.py:42,123-126).
    One slice against the fp32 oracle on the host."""
    from oracle import sam_fp32
    from vit_deep_radiomics_b200 import tfds_dense_descriptor as tdd
    model = tdd.load_model("medsam", device=cuda, seed=17)
    assert model.img_hw == (1024, 1024) and model.grid == (64, 64) and model.feature_dim == 256
    gray = torch.rand(1, 1024, 1024, generator=torch.Generator().manual_seed(5))
    got = model.dense_descriptors(gray.to(cuda)).cpu().numpy()
    torch.set_num_threads(max(1, os.cpu_count() or 1))
    with torch.no_grad():
        want = sam_fp32.sam_dense_descriptor(model.state_dict_f32, model.cfg, gray[:, None].expand(-1, 3, -1, -1)).numpy()
    assert got.shape == want.shape == (1, 64, 64, 256)
    _check_descriptors(got, want, "descriptors medsam (SAM ViT-B)@1024x1024")


def test_medsam_point_cloud_through_the_gather(cuda):
    """extract_point_cloud with the SAM encoder (no CLS row: token offset 0): indices bit-exact, tokens in tolerance."""
    from oracle import gather_np, sam_fp32
    from vit_deep_radiomics_b200 import tfds_dense_descriptor as tdd
    rng = np.random.default_rng(8)
    H = W = 128
    S = 3
    img = rng.random((H, W, S), dtype=np.float32)
    yy, xx = np.mgrid[:H, :W]
    mask = np.stack([((yy - 64) ** 2 + (xx - 60) ** 2) < (14 + 2 * s) ** 2 for s in range(S)], axis=-1)
    res = np.array([0.8, 0.8, 0.8])
    model = tdd.load_model("sam_tiny", img_hw=(256, 256), device=cuda, seed=19)
    out = tdd.extract_point_cloud(model, img, mask, res)
    y0, y1, x0, x1 = out["plan"]["crop"]
    # oracle: the crop window resized by the product's own device resize (its parity is tests/test_gpu_pipeline.py's subject),
    # then the fp32 encoder and the reference-order gather
    from vit_deep_radiomics_b200 import ops
    sl = ops.volume_to_slices(torch.from_numpy(img).to(cuda), out["plan"]["crop"], out_hw=model.img_hw).float().cpu()
    with torch.no_grad():
        dense = sam_fp32.sam_dense_descriptor(model.state_dict_f32, model.cfg, sl[:, None].expand(-1, 3, -1, -1)).numpy()
    fy0, fy1, fx0, fx1 = out["plan"]["feat_roi"]
    my0, my1, mx0, mx1 = out["plan"]["mask_roi"]
    mask_c = mask[y0:y1, x0:x1]
    ref = gather_np.token_gather([dense[s, fy0:fy1, fx0:fx1] for s in range(S)], [mask_c[my0:my1, mx0:mx1, s] for s in range(S)], res)
    assert out["count"] == ref["flat"].size and out["count"] > 0
    assert np.array_equal(out["src"].numpy(), ref["src"])
    got, want = out["tokens"].numpy().astype(np.float64), ref["tokens"]
    cos = (got * want).sum(1) / (np.linalg.norm(got, axis=1) * np.linalg.norm(want, axis=1))
    assert np.abs(got - want).max() < ABS_TOL and cos.min() > COS


def test_medsam_descriptors_feed_the_shipped_classifier_config(cuda):
    """The reference's default pipeline end to end: SAM-encoder descriptors (256 channels, conf/parameters_models.yaml
    feature_dim: 256) -> tumour-mask gather -> TransformerNoduleClassifier built by build_model from the shipped YAML ->
    focal loss -> backward.  Logits / CLS / loss on the extracted tokens against the fp32 oracle classifier on the same tokens."""
    import os
    from oracle import classifier_fp32 as C
    from vit_deep_radiomics_b200 import config_manager, tfds_dense_descriptor as tdd, train_models as tm
    rng = np.random.default_rng(4)
    H = W = 128
    S = 4
    img = rng.random((H, W, S), dtype=np.float32)
    yy, xx = np.mgrid[:H, :W]
    mask = np.stack([((yy - 60) ** 2 + (xx - 66) ** 2) < (12 + 2 * s) ** 2 for s in range(S)], axis=-1)
    model = tdd.load_model("sam_small", img_hw=(256, 256), device=cuda, seed=31)
    assert model.feature_dim == 256
    out = tdd.extract_point_cloud(model, img, mask, np.array([0.8, 0.8, 0.8]), to_host=False)
    n = int(out["count"].item())
    assert n > 0
    tokens = out["tokens"][:n]
    cfg = config_manager.load_conf(project_dir=os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
    torch.manual_seed(2)
    from vit_deep_radiomics_b200.models_archs import set_dropout
    clf = set_dropout(tm.build_model(cfg, "transformer", "ct").to(cuda), 0.0, 0.0)   # compared with the dropout-free oracle below
    mcfg = cfg["models"]["transformer"]
    assert mcfg["feature_dim"] == 256
    heads, layers = mcfg["ct"]["num_heads"], mcfg["ct"]["num_layers"]
    y = torch.tensor([0.0, 1.0], device=cuda)
    logits, cls = clf(tokens.unsqueeze(0))
    loss = tm.FocalLoss(alpha=torch.tensor([0.25, 0.75], device=cuda), gamma=2)(torch.squeeze(logits), y)
    loss.backward()
    sd = {k: v.detach().cpu().clone() for k, v in clf.state_dict().items()}
    lg, cl = C.classifier_forward(sd, tokens.detach().cpu()[None], heads, layers)
    ref_loss = C.focal_loss(lg[0], y.cpu(), 2.0, torch.tensor([0.25, 0.75]))
    assert torch.allclose(logits.detach().cpu().reshape(-1), lg.reshape(-1), atol=0.05, rtol=0.05)
    assert F.cosine_similarity(cls.detach().cpu().reshape(1, -1), cl.reshape(1, -1)).item() > 0.999
    assert abs(float(loss) - float(ref_loss)) < 0.05 * max(1.0, abs(float(ref_loss)))
    for k, p in clf.named_parameters():
        assert p.grad is not None and torch.isfinite(p.grad).all(), k


@pytest.mark.parametrize("hw,B", [((56, 84), 2), ((896, 896), 1)])
def test_dinov2_patch_embed_mode(cuda, hw, B):
    """The reference's 'dinov2' branch only calls model.patch_embed (:128-133)."""
    from oracle import sam_fp32
    from vit_deep_radiomics_b200 import tfds_dense_descriptor as tdd
    model = tdd.load_model("dinov2", img_hw=hw, device=cuda, seed=23)
    x = torch.rand(B, 3, *hw, generator=torch.Generator().manual_seed(6))
    got = model.dense_descriptors(x.to(cuda)).cpu()
    want = sam_fp32.dinov2_patch_embed(model.state_dict_f32, x)
    assert got.shape == want.shape == (B, hw[0] // 14, hw[1] // 14, 384)
    assert torch.allclose(got, want, atol=6e-3, rtol=1e-2), float((got - want).abs().max())


@pytest.mark.parametrize("hw,B", [((256, 256), 2), ((224, 224), 1), ((128, 1024), 1)])
def test_sam_native_forward_equals_op_by_op(cuda, hw, B):
    """vdr_sam_forward (the whole encoder as one C call) enqueues the same kernels in the same order as the op-by-op path:
    bit-identical descriptors from a materialised im2col matrix (RGB pictures); gray slices of a volume go through the TMA im2col
    patch embedding on channel-summed weights where the geometry allows it (another rounding of the same sum: tolerance)."""
    from vit_deep_radiomics_b200 import _C, sam_encoder
    model = sam_encoder.SamImageEncoder("sam_tiny", img_hw=hw, device=cuda, seed=31)
    x = torch.rand(B, 3, *hw, generator=torch.Generator().manual_seed(6)).to(cuda)
    assert model._native_struct() is not None
    n0 = _C.launch_count()
    native = model.dense_descriptors(x).clone()
    n_native = _C.launch_count() - n0
    model.use_native_forward = False
    assert model._native_struct() is None
    eager = model.dense_descriptors(x).clone()
    assert n_native > 0 and torch.equal(native, eager)
    # a gray volume: (H, W, S) f32, full-window crop
    vol = torch.rand(hw[0], hw[1], 3, generator=torch.Generator().manual_seed(7)).to(cuda)
    eager_v = model.forward_volume(vol, (0, hw[0], 0, hw[1])).clone()
    model.use_native_forward = True
    native_v = model.forward_volume(vol, (0, hw[0], 0, hw[1])).clone()
    assert native_v.shape == eager_v.shape == (3 * model.n_tokens, model.feature_dim)
    err = (native_v - eager_v).abs().max().item()
    assert err <= 0.05, err                                     # O(1) LayerNorm outputs; bf16 patch weights summed vs summed products
    cos = torch.nn.functional.cosine_similarity(native_v, eager_v, dim=1).min().item()
    assert cos >= 0.9995, cos


@pytest.mark.parametrize("gain,nats", [(4.0, 30), (14.0, 120)])
def test_fused_relpos_attention_maximum_free_blocks_guard(cuda, gain, nats):
    """The fused global-attention kernel takes its reference maximum from key block 0 (two rows of the token grid) and runs the later
    blocks without a maximum.  Keys further down whose scores sit tens of nats above it (large P, no overflow) or > 88 nats above it
    (exp2 overflows: the row sums flag the tile and the CTA recomputes it exactly, bias included, on the CUDA cores) must give the
    fp32 softmax either way."""
    from vit_deep_radiomics_b200 import ops
    g = torch.Generator().manual_seed(int(gain))
    BW, Sh, Sw, heads = 1, 8, 64, 2
    N, d = Sh * Sw, heads * 64
    qkv = torch.randn(BW * N, 3 * d, generator=g)
    x = qkv.view(BW, N, 3, heads, 64)
    x[:, :, 0] = x[:, :, 0].abs() * 0.5 + 2.0
    x[:, :, 1] = x[:, :, 1].abs() * 0.1 + 0.5
    x[:, 200:330, 1] *= gain                                    # rows 3 .. 5 of the grid: key blocks 1 and 2
    qkv = qkv.bfloat16()
    rel_h = torch.randn(2 * Sh - 1, 64, generator=g) * 0.1
    rel_w = torch.randn(2 * Sw - 1, 64, generator=g) * 0.1
    want, _ = _attn_reference(qkv, BW, Sh, Sw, heads, rel_h, rel_w)
    q, k, _ = qkv.float().reshape(BW, N, 3, heads, 64).permute(2, 0, 3, 1, 4)
    s = q @ k.transpose(-1, -2) / 8.0
    assert (s[..., 128:].amax(-1) - s[..., :128].amax(-1)).max().item() > nats
    hi, lo = ops.relpos_split(rel_h.to(cuda), rel_w.to(cuda))
    fused = ops.attn_relpos(qkv.to(cuda), BW, Sh, Sw, heads, hi, lo, kernel="fused").cpu().float()
    assert torch.isfinite(fused).all()
    err = (fused - want).abs().max()
    cos = F.cosine_similarity(fused.reshape(-1, 64), want.reshape(-1, 64), dim=1).min()
    assert err < 2e-2 * max(1.0, float(want.abs().max()) / 3.0) and cos > 0.9995, (float(err), float(cos))
