"""-m gpu: each CUDA kernel (through the C ABI) against a plain PyTorch fp32 reference of the same op.
Tolerances are for bf16 operands / bf16 outputs with fp32 accumulation (bf16 eps = 2^-8 = 3.9e-3)."""
import math

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


def _close(got, ref, rel_to_peak):
    got, ref = got.double().cpu(), ref.double().cpu()
    assert torch.isfinite(got).all()
    err = (got - ref).abs().max().item()
    assert err <= rel_to_peak * max(ref.abs().max().item(), 1e-6), f"max abs err {err} vs peak {ref.abs().max().item()}"


@pytest.mark.parametrize("M,N,K,epi,cdt", [
    (128, 256, 64, "bias", torch.float32), (300, 384, 200, "gelu", torch.bfloat16),
    (1025, 768, 768, "residual", torch.bfloat16), (4100, 2304, 768, "bias", torch.bfloat16),
    (1, 512, 256, "gelu", torch.bfloat16), (130, 8, 72, "bias", torch.float32),
    (1576, 1152, 384, "bias", torch.bfloat16), (257, 1024, 592, "residual", torch.float32),
    (2050, 768, 3072, "residual", torch.bfloat16)])
def test_gemm(cuda, M, N, K, epi, cdt):
    from vit_deep_radiomics_b200 import ops
    torch.manual_seed(M + N + K)
    a = (torch.randn(M, K, device=cuda) * 0.5).bfloat16()
    w = (torch.randn(N, K, device=cuda) * 0.05).bfloat16()
    b = torch.randn(N, device=cuda) * 0.1
    r = torch.randn(M, N, device=cuda).to(cdt) if epi == "residual" else None
    ref = a.float() @ w.float().t() + b
    if epi == "gelu":
        ref = torch.nn.functional.gelu(ref)
    if epi == "residual":
        ref = ref + r.float()
    out = ops.gemm(a, w, b, epilogue=epi, residual=r, out_dtype=cdt)
    _close(out, ref, 1e-5 if cdt == torch.float32 else 5e-3)


@pytest.mark.parametrize("M,d,N,epi", [(1025, 768, 2304, "bias"), (4100, 768, 3072, "gelu"), (333, 384, 1152, "bias"),
                                       (40000, 768, 2304, "bias"), (70, 192, 576, "gelu"), (2050, 1024, 4096, "gelu")])
def test_gemm_folded_layernorm_consumer(cuda, M, d, N, epi):
    """`x -> LayerNorm -> Linear` as ONE GEMM on the raw x (vdr_gemm ln_stats / ln_colsum + vdr_fold_layernorm + vdr_row_stats)
    against fp32 LayerNorm + matmul; x carries a per-row offset and scale so mean and rstd both matter."""
    from vit_deep_radiomics_b200 import ops
    torch.manual_seed(M + N)
    x = ((torch.randn(M, d, device=cuda) * (0.5 + 2 * torch.rand(M, 1, device=cuda)) + torch.randn(M, 1, device=cuda))).bfloat16()
    w = (torch.randn(N, d, device=cuda) * 0.05).bfloat16()
    b = torch.randn(N, device=cuda) * 0.1
    gamma, beta = 1 + 0.3 * torch.randn(d, device=cuda), 0.2 * torch.randn(d, device=cuda)
    ref = torch.nn.functional.layer_norm(x.float(), (d,), gamma, beta, eps=1e-6) @ w.float().t() + b
    if epi == "gelu":
        ref = torch.nn.functional.gelu(ref)
    wf, bf, cs = ops.fold_layernorm(w, b, gamma, beta)
    assert torch.equal(wf, (w.float() * gamma).bfloat16())
    assert torch.allclose(cs, wf.float().sum(1), atol=1e-4) and torch.allclose(bf, b + w.float() @ beta, atol=1e-4)
    st = ops.row_stats(x)
    assert st.shape == (1, M, 2)
    assert torch.allclose(st[0, :, 0], x.float().sum(1), rtol=1e-5, atol=1e-3)
    assert torch.allclose(st[0, :, 1], (x.float() ** 2).sum(1), rtol=1e-5, atol=1e-3)
    out = ops.gemm(x, wf, bf, epilogue=epi, ln_stats=st, ln_colsum=cs, ln_eps=1e-6)
    _close(out, ref, 6e-3)
    # the same statistics split over several slots (the layout a residual GEMM's stats_out produces) give the same result
    st3 = torch.stack([st[0] * 0.25, st[0] * 0.5, st[0] * 0.25]).contiguous()
    out3 = ops.gemm(x, wf, bf, epilogue=epi, ln_stats=st3, ln_colsum=cs, ln_eps=1e-6)
    _close(out3, out.float(), 1e-2)


@pytest.mark.parametrize("M,N,K", [(1025, 768, 768), (2050, 768, 3072), (40000, 768, 768), (300, 384, 1536), (90, 192, 192), (130, 1024, 256)])
def test_gemm_residual_emits_row_statistics(cuda, M, N, K):
    """stats_out of a residual GEMM: per row and 64-column slot, (sum, sum of squares) of the values written (in place on the
    residual, as the ViT blocks run it); deterministic, and feeding it to a folded consumer equals LayerNorm of the output."""
    from vit_deep_radiomics_b200 import ops
    torch.manual_seed(M + K)
    a = (torch.randn(M, K, device=cuda) * 0.5).bfloat16()
    w = (torch.randn(N, K, device=cuda) * 0.05).bfloat16()
    b = torch.randn(N, device=cuda) * 0.1
    x0 = torch.randn(M, N, device=cuda).bfloat16()
    outs = []
    for _ in range(2):
        x = x0.clone()
        st = torch.full((N // 64, M, 2), float("nan"), device=cuda)
        ops.gemm(a, w, b, epilogue="residual", residual=x, out=x, stats_out=st)
        outs.append((x, st))
    assert torch.equal(outs[0][0], outs[1][0]) and torch.equal(outs[0][1], outs[1][1])
    x, st = outs[0]
    _close(x, a.float() @ w.float().t() + b + x0.float(), 5e-3)
    assert torch.equal(x, ops.gemm(a, w, b, epilogue="residual", residual=x0))          # statistics do not change the output
    xs = x.float().reshape(M, N // 64, 64)
    # statistics are taken before the bf16 rounding of the output: compare within that rounding
    assert torch.allclose(st[:, :, 0].t(), xs.sum(2), atol=0.15, rtol=0)
    assert torch.allclose(st[:, :, 1].t(), (xs ** 2).sum(2), rtol=1e-2, atol=0.05)
    d = N
    w2 = (torch.randn(256, d, device=cuda) * 0.05).bfloat16()
    gamma, beta = 1 + 0.3 * torch.randn(d, device=cuda), 0.2 * torch.randn(d, device=cuda)
    wf, bf, cs = ops.fold_layernorm(w2, None, gamma, beta)
    out = ops.gemm(x, wf, bf, ln_stats=st, ln_colsum=cs, ln_eps=1e-6)
    ref = torch.nn.functional.layer_norm(x.float(), (d,), gamma, beta, eps=1e-6) @ w2.float().t()
    _close(out, ref, 6e-3)


def test_gemm_inplace_residual_tile_reuse_hazard(cuda):
    """Regression: the TMA-loaded residual tile of an epilogue warp is overwritten by the next chunk's load; without a fence after
    the (asynchronously completing) shared-memory reads, a read delayed behind the main loop's traffic saw the next chunk's bytes --
    a 16-byte corruption in a few rows, a few times per hundred launches (pair tiles, K = 768).  40 launches must agree bit for bit."""
    from vit_deep_radiomics_b200 import ops
    torch.manual_seed(5)
    M, N, K = 60000, 768, 768
    a = (torch.randn(M, K, device=cuda) * 0.5).bfloat16()
    w = (torch.randn(N, K, device=cuda) * 0.05).bfloat16()
    b = torch.randn(N, device=cuda) * 0.1
    x0 = torch.randn(M, N, device=cuda).bfloat16()
    ref = None
    for _ in range(40):
        x = x0.clone()
        ops.gemm(a, w, b, epilogue="residual", residual=x, out=x)
        if ref is None:
            ref = x
            _close(x, a.float() @ w.float().t() + b + x0.float(), 5e-3)
        else:
            assert torch.equal(x, ref)


def test_gemm_folded_layernorm_argument_errors(cuda):
    from vit_deep_radiomics_b200 import ops
    a = torch.zeros(64, 128, device=cuda, dtype=torch.bfloat16)
    w = torch.zeros(96, 128, device=cuda, dtype=torch.bfloat16)
    st = torch.zeros(1, 64, 2, device=cuda)
    with pytest.raises(ValueError):          # stats_out needs the residual epilogue and N % 64 == 0
        ops.gemm(a, w, None, stats_out=torch.zeros(2, 64, 2, device=cuda))
    with pytest.raises(ValueError):          # f32 output with a folded LayerNorm
        ops.gemm(a, w, None, ln_stats=st, ln_colsum=torch.zeros(96, device=cuda), out_dtype=torch.float32)
    with pytest.raises(ValueError):
        ops.gemm(a, w, None, ln_stats=torch.zeros(1, 63, 2, device=cuda), ln_colsum=torch.zeros(96, device=cuda))


def test_gemm_gelu_epilogue_accuracy_over_the_whole_range(cuda):
    """The bf16 epilogue's GELU (x * sigmoid of an odd quintic) against the erf form: identity weights so
    that every output is gelu(a[m, n]); inputs sweep [-12, 12] plus saturated values.  Bound: 3e-5 absolute + one bf16 rounding."""
    from vit_deep_radiomics_b200 import ops
    M, N = 4096, 64
    a = torch.linspace(-12, 12, M * N, device=cuda).reshape(M, N)
    a[0, :8] = torch.tensor([-100.0, -30.0, -16.0, 16.0, 30.0, 100.0, 0.0, -0.0], device=cuda)
    a = a.bfloat16()
    w = torch.eye(N, device=cuda).bfloat16()
    out = ops.gemm(a, w, None, epilogue="gelu").double()
    ref = torch.nn.functional.gelu(a.double())
    assert torch.isfinite(out).all()
    err = (out - ref).abs()
    assert (err <= 3e-5 + ref.abs() * 2.0 ** -8).all(), float(err.max())
    assert float(out[0, 0]) == 0.0 and float(out[0, 5]) == 100.0 and float(out[0, 4]) == 30.0


def test_gemm_k_tail_and_row_remap(cuda):
    """K = 588 (14x14 patches) inside ld 592, rows written behind a CLS row, pos-embed as residual."""
    from vit_deep_radiomics_b200 import ops
    torch.manual_seed(0)
    B, Np, d, K, ld = 3, 256, 1024, 588, 592
    a = torch.zeros(B * Np, ld, device=cuda, dtype=torch.bfloat16)
    a[:, :K] = (torch.randn(B * Np, K, device=cuda) * 0.5).bfloat16()
    a[:, K:] = 7.0                                              # garbage in the padding must be ignored
    w = torch.zeros(d, ld, device=cuda, dtype=torch.bfloat16)
    w[:, :K] = (torch.randn(d, K, device=cuda) * 0.05).bfloat16()
    w[:, K:] = -3.0
    bias, pos = torch.randn(d, device=cuda) * 0.1, torch.randn(Np + 1, d, device=cuda)
    out = torch.zeros(B * (Np + 1), d, device=cuda, dtype=torch.bfloat16)
    ops.gemm(a, w, bias, epilogue="residual", residual=pos, out=out, k=K, out_group=(Np, Np + 1, 1), res_mod=(Np, 1))
    ref = (a[:, :K].float() @ w[:, :K].float().t() + bias).reshape(B, Np, d) + pos[1:]
    o3 = out.reshape(B, Np + 1, d)
    _close(o3[:, 1:], ref, 5e-3)
    assert (o3[:, 0] == 0).all()


def test_gemm_argument_errors(cuda):
    from vit_deep_radiomics_b200 import ops
    a = torch.zeros(16, 64, device=cuda, dtype=torch.bfloat16)
    with pytest.raises(ValueError):
        ops.gemm(a, torch.zeros(12, 64, device=cuda, dtype=torch.bfloat16))          # N % 8 != 0
    with pytest.raises(ValueError):
        ops.gemm(a[:, 1:], torch.zeros(16, 63, device=cuda, dtype=torch.bfloat16))   # misaligned / ld % 8


@pytest.mark.parametrize("B,N,heads", [(1, 128, 1), (2, 197, 6), (1, 1025, 2), (3, 300, 4), (1, 7, 1), (2, 257, 16),
                                         (2, 136, 3), (2, 1030, 2), (3, 645, 2), (2, 2, 1), (5, 1025, 12), (2, 250, 2), (3, 1273, 1)])   # 8 / 6 / 5 / 2 trailing rows: the mma.sync tail CTA
def test_flash_attention(cuda, B, N, heads):
    from vit_deep_radiomics_b200 import ops
    torch.manual_seed(B * 1000 + N)
    d = heads * 64
    qkv = torch.randn(B * N, 3 * d, device=cuda).bfloat16()
    q, k, v = qkv.float().reshape(B, N, 3, heads, 64).permute(2, 0, 3, 1, 4)
    s = q @ k.transpose(-1, -2) / math.sqrt(64)
    ref = (torch.softmax(s, -1) @ v).transpose(1, 2).reshape(B * N, d)
    out, lse = ops.flash_attn(qkv, B, N, heads, return_lse=True)
    _close(out, ref, 8e-3)
    _close(lse, torch.logsumexp(s, -1), 5e-4)


@pytest.mark.parametrize("N,late_gain,nats", [(1025, 5.0, 35), (1025, 16.0, 120), (645, 16.0, 120), (384, 16.0, 120)])
def test_flash_attention_maximum_free_blocks_guard(cuda, N, late_gain, nats):
    """The plain kernel takes its reference maximum from key block 0 only.  Later keys whose scores sit tens of nats above it give
    P >> 1 (no overflow: same result as the exact softmax); hundreds of nats above it overflow exp2, the row sums flag it and the
    CTA recomputes its tile exactly (attn_tail_rows) -- outputs and lse must equal the fp32 softmax either way."""
    from vit_deep_radiomics_b200 import ops
    torch.manual_seed(N + int(late_gain * 10))
    B, heads = 2, 2
    d = heads * 64
    qkv = torch.randn(B * N, 3 * d, device=cuda)
    x = qkv.view(B, N, 3, heads, 64)
    x[:, :, 0] = x[:, :, 0].abs() * 0.5 + 2.0                   # queries: all components in [2, ~4]
    x[:, :, 1] = x[:, :, 1].abs() * 0.1 + 0.5                   # keys: positive, so q . k grows with the key's gain
    x[:, 200:, 1] *= late_gain                                   # keys beyond block 0 (and a part of block 1) are scaled
    if N > 600:
        x[0, 600:, 1] = x[0, 600:, 1] / late_gain                # image 0: only blocks 1 .. 4 carry the large keys
    qkv = qkv.bfloat16()
    q, k, v = qkv.float().reshape(B, N, 3, heads, 64).permute(2, 0, 3, 1, 4)
    s = q @ k.transpose(-1, -2) / math.sqrt(64)
    gap = (s[..., 128:].amax(-1) - s[..., :128].amax(-1))        # nats above (below) the block-0 maximum, per row
    assert gap.max().item() > nats
    ref = (torch.softmax(s, -1) @ v).transpose(1, 2).reshape(B * N, d)
    out, lse = ops.flash_attn(qkv, B, N, heads, return_lse=True)
    assert torch.isfinite(out.float()).all() and torch.isfinite(lse).all()
    _close(out, ref, 8e-3)
    _close(lse, torch.logsumexp(s, -1), 5e-4)


@pytest.mark.parametrize("hot_key", [130, 131, 138, 129, 600, 1024])
def test_flash_attention_single_hot_key_beyond_block0(cuda, hot_key):
    """One key with a score hundreds of nats above everything else, sitting in a column that takes the polynomial exp2 (130, 131, 138),
    a MUFU column (129), a later block or the trailing key (1024): a maximum-free block must not turn it into a wrapped exponent."""
    from vit_deep_radiomics_b200 import ops
    torch.manual_seed(hot_key)
    B, N, heads = 1, 1025, 1
    qkv = torch.randn(B * N, 3 * 64, device=cuda)
    qkv[:, :64] = qkv[:, :64].abs() + 1.0                       # queries positive
    qkv[:, 64:128] *= 0.1
    qkv[hot_key, 64:128] = 24.0                                  # q . k / 8 ~ 64 * 1.8 * 24 / 8 = 350 nats
    qkv = qkv.bfloat16()
    q, k, v = qkv.float().reshape(B, N, 3, heads, 64).permute(2, 0, 3, 1, 4)
    s = q @ k.transpose(-1, -2) / math.sqrt(64)
    assert (s[..., hot_key] - s[..., :128].amax(-1)).min().item() > 150
    ref = (torch.softmax(s, -1) @ v).transpose(1, 2).reshape(B * N, 64)
    out, lse = ops.flash_attn(qkv, B, N, heads, return_lse=True)
    _close(out, ref, 8e-3)
    _close(lse, torch.logsumexp(s, -1), 5e-4)


@pytest.mark.parametrize("rows,d", [(1000, 256), (1025, 768), (333, 384), (77, 1024), (50, 2048), (3, 8)])
def test_layernorm_fwd_bwd(cuda, rows, d):
    from vit_deep_radiomics_b200 import ops
    torch.manual_seed(rows + d)
    x = (torch.randn(rows, d, device=cuda) * 2 + 0.5).bfloat16()
    g, b = torch.randn(d, device=cuda), torch.randn(d, device=cuda)
    ref = torch.nn.functional.layer_norm(x.float(), (d,), g, b, 1e-6)
    _close(ops.layernorm(x, g, b, 1e-6, out_dtype=torch.float32), ref, 1e-6)
    y, mu, rs = ops.layernorm(x, g, b, 1e-6, save_stats=True)
    _close(y, ref, 5e-3)
    _close(mu, x.float().mean(1), 1e-5)
    if d <= 1024:
        dy = torch.randn(rows, d, device=cuda).bfloat16()
        xr, gr, br = x.float().requires_grad_(True), g.clone().requires_grad_(True), b.clone().requires_grad_(True)
        torch.nn.functional.layer_norm(xr, (d,), gr, br, 1e-6).backward(dy.float())
        dg, db = torch.zeros(d, device=cuda), torch.zeros(d, device=cuda)
        dx = ops.layernorm_bwd(dy, x, g, mu, rs, dg, db)
        _close(dx, xr.grad, 5e-3)
        _close(dg, gr.grad, 1e-4)
        _close(db, br.grad, 1e-4)


def test_cls_concat_layernorm(cuda):
    from vit_deep_radiomics_b200 import ops
    torch.manual_seed(1)
    for n, d in [(777, 256), (0, 256), (5, 768)]:
        x, cls = torch.randn(n, d, device=cuda), torch.randn(d, device=cuda)
        g, b = torch.randn(d, device=cuda), torch.randn(d, device=cuda)
        ref = torch.nn.functional.layer_norm(torch.cat([cls[None], x]), (d,), g, b, 1e-5)
        _close(ops.cls_concat_layernorm(x, cls, g, b, 1e-5), ref, 5e-3)


def test_im2col_bit_exact(cuda):
    from vit_deep_radiomics_b200 import ops
    torch.manual_seed(5)
    for (B, H, W, p, gray) in [(2, 64, 64, 16, True), (2, 56, 42, 14, False), (3, 224, 224, 16, True)]:
        if gray:   # (H, W, S) volume layout as np.dstack gives; gray2rgb by channel stride 0
            vol = torch.randn(H, W, B, device=cuda)
            img = vol.permute(2, 0, 1)[:, None].expand(B, 3, H, W)
            A = ops.im2col_patches(vol, (1, 0, W * B, B), B, H, W, p)
        else:
            img = torch.randn(B, 3, H, W, device=cuda)
            A = ops.im2col_patches(img, img.stride(), B, H, W, p)
        ref = torch.nn.functional.unfold(img.contiguous(), kernel_size=p, stride=p).transpose(1, 2).reshape(-1, 3 * p * p)
        K = 3 * p * p
        assert torch.equal(A[:, :K].float(), ref.bfloat16().float()) and (A[:, K:] == 0).all()


def test_volume_staging_and_gray_im2col(cuda):
    """(H, W, S) f32 volume -> crop -> (S, ch, cw) bf16 slices -> im2col with replicated channels
    == unfold of gray2rgb(crop) (prepare_image + patch-embed data movement), bit-exact in bf16."""
    from vit_deep_radiomics_b200 import ops
    torch.manual_seed(6)
    for (H, W, S, crop, p) in [(64, 64, 5, (0, 64, 0, 64), 16), (80, 72, 37, (8, 72, 4, 68), 16), (60, 70, 3, (2, 58, 0, 70), 14)]:
        vol = torch.rand(H, W, S, device=cuda)
        y0, y1, x0, x1 = crop
        sl = ops.volume_to_slices(vol, crop)
        want = vol[y0:y1, x0:x1].permute(2, 0, 1).contiguous()
        assert torch.equal(sl.float(), want.bfloat16().float())
        A = ops.im2col_gray_bf16(sl, p)
        ref = torch.nn.functional.unfold(want[:, None].expand(-1, 3, -1, -1).contiguous(), kernel_size=p, stride=p)
        ref = ref.transpose(1, 2).reshape(-1, 3 * p * p)
        K = 3 * p * p
        assert torch.equal(A[:, :K].float(), ref.bfloat16().float()) and (A[:, K:] == 0).all()


@pytest.mark.parametrize("B,C,H,W,d", [(3, 1, 512, 512, 768), (40, 1, 512, 512, 768), (2, 3, 256, 256, 384), (1, 1, 128, 1024, 128)])
def test_patch_embed_tma_im2col_matches_materialised_path(cuda, B, C, H, W, d):
    """vdr_patch_embed_gemm (5-D TMA im2col view, no A matrix) == vdr_im2col_patches + vdr_gemm: same K order, same
    accumulation order -> bit-identical bf16 tokens; CLS rows are left untouched."""
    from vit_deep_radiomics_b200 import ops
    torch.manual_seed(3)
    p = 16
    assert ops.patch_embed_supported(H, W, p)
    imgs = torch.rand((B, H, W) if C == 1 else (B, 3, H, W), device=cuda).bfloat16()
    gh, gw = H // p, W // p
    Np, N, K = gh * gw, gh * gw + 1, 3 * p * p
    w_pe = (torch.randn(d, K, device=cuda) * 0.05).bfloat16()
    bias, pos = torch.randn(d, device=cuda), torch.randn(N, d, device=cuda)
    x_ref = torch.full((B * N, d), 7.0, device=cuda, dtype=torch.bfloat16)
    src = imgs.float()
    strides = (src.stride(0), 0, src.stride(1), src.stride(2)) if C == 1 else src.stride()
    A = torch.empty(B * Np, K, device=cuda, dtype=torch.bfloat16)
    ops.im2col_patches(src, strides, B, H, W, p, out=A)
    ops.gemm(A, w_pe, bias, epilogue="residual", residual=pos, out=x_ref, k=K, out_group=(Np, N, 1), res_mod=(Np, 1))
    x = torch.full((B * N, d), 7.0, device=cuda, dtype=torch.bfloat16)
    ops.patch_embed(imgs, w_pe, bias, pos, p, out=x)
    assert torch.equal(x, x_ref)
    assert torch.all(x.view(B, N, d)[:, 0] == 7.0)
    # against plain fp32 arithmetic (tolerance of bf16 operands / outputs)
    patches = src.reshape(B, -1, gh, p, gw, p) if C == 3 else src[:, None].expand(-1, 3, -1, -1).reshape(B, 3, gh, p, gw, p)
    a32 = patches.permute(0, 2, 4, 1, 3, 5).reshape(B * Np, K)
    want = a32 @ w_pe.float().t() + bias + pos[1:].repeat(B, 1)
    got = x.view(B, N, d)[:, 1:].reshape(B * Np, d).float()
    assert (got - want).abs().max() <= 0.02 * want.abs().max()


@pytest.mark.parametrize("B,H,W,d", [(3, 512, 512, 768), (2, 256, 256, 384)])
def test_patch_embed_gray_channel_summed_weights(cuda, B, H, W, d):
    """vdr_patch_embed_gemm_gray: gray pictures against W_r + W_g + W_b (K = p*p) == the exact fp32 patch embedding of the
    gray2rgb picture within the rounding of the summed bf16 weights, and agrees with the 3-channel TMA path."""
    from vit_deep_radiomics_b200 import ops
    torch.manual_seed(5)
    p = 16
    imgs = torch.rand((B, H, W), device=cuda).bfloat16()
    gh, gw = H // p, W // p
    Np, N = gh * gw, gh * gw + 1
    w4 = torch.randn(d, 3, p, p, device=cuda) * 0.05
    w_sum = w4.sum(dim=1).reshape(d, p * p).bfloat16().contiguous()
    bias, pos = torch.randn(d, device=cuda), torch.randn(N, d, device=cuda)
    x = torch.full((B * N, d), 7.0, device=cuda, dtype=torch.bfloat16)
    ops.patch_embed(imgs, w_sum, bias, pos, p, out=x, channel_summed=True)
    assert torch.all(x.view(B, N, d)[:, 0] == 7.0)
    a32 = imgs.float().reshape(B, gh, p, gw, p).permute(0, 1, 3, 2, 4).reshape(B * Np, p * p)
    got = x.view(B, N, d)[:, 1:].reshape(B * Np, d).float()
    want_sum = a32 @ w_sum.float().t() + bias + pos[1:].repeat(B, 1)          # same bf16 operands: only the output rounding differs
    assert (got - want_sum).abs().max() <= 0.006 * want_sum.abs().max()
    want = a32 @ w4.sum(dim=1).reshape(d, -1).t() + bias + pos[1:].repeat(B, 1)   # exact weights
    assert (got - want).abs().max() <= 0.02 * want.abs().max()
    x3 = torch.full((B * N, d), 7.0, device=cuda, dtype=torch.bfloat16)
    ops.patch_embed(imgs, w4.reshape(d, -1).bfloat16().contiguous(), bias, pos, p, out=x3)
    got3 = x3.view(B, N, d)[:, 1:].reshape(B * Np, d).float()
    assert (got - got3).abs().max() <= 0.02 * want.abs().max()
    with pytest.raises(ValueError):
        ops.patch_embed(imgs[:, None].expand(-1, 3, -1, -1).contiguous(), w_sum, bias, pos, p, out=x, channel_summed=True)


def test_patch_embed_unsupported_geometry_is_rejected(cuda):
    from vit_deep_radiomics_b200 import ops
    assert not ops.patch_embed_supported(224, 224, 16)      # 14 patches per row do not tile a 128-row box
    assert not ops.patch_embed_supported(518, 518, 14)
    imgs = torch.zeros(1, 224, 224, device=cuda, dtype=torch.bfloat16)
    with pytest.raises(ValueError):
        ops.patch_embed(imgs, torch.zeros(64, 768, device=cuda, dtype=torch.bfloat16), torch.zeros(64, device=cuda),
                        torch.zeros(197, 64, device=cuda), 16, out=torch.zeros(197, 64, device=cuda, dtype=torch.bfloat16))


@pytest.mark.parametrize("B,N,heads", [(1, 128, 1), (1, 300, 2), (2, 257, 4), (1, 1025, 2), (1, 2500, 4)])
def test_flash_attention_backward_vs_fp32_autograd(cuda, B, N, heads):
    """vdr_flash_attn_bwd against torch fp32 autograd of softmax(q k^T / 8) v on the same bf16 inputs: dq, dk, dv within bf16
    operand tolerance (P and dS are rounded to bf16 for the tensor cores, as in the forward)."""
    from vit_deep_radiomics_b200 import ops
    torch.manual_seed(N + heads)
    d = heads * 64
    qkv = (torch.randn(B * N, 3 * d, device=cuda) * 0.8).bfloat16()
    do = (torch.randn(B * N, d, device=cuda) * 0.5).bfloat16()
    out, lse = ops.flash_attn(qkv, B, N, heads, return_lse=True)
    dqkv = ops.flash_attn_bwd(qkv, out, do, lse, B, N, heads).float()
    x = qkv.float().view(B, N, 3, heads, 64).permute(2, 0, 3, 1, 4).contiguous().requires_grad_(True)     # (3, B, h, N, 64)
    q, k, v = x[0], x[1], x[2]
    p = torch.softmax(q @ k.transpose(-1, -2) / 8.0, dim=-1)
    o = p @ v
    o.backward(do.float().view(B, N, heads, 64).permute(0, 2, 1, 3))
    want = x.grad.permute(1, 3, 0, 2, 4).reshape(B * N, 3 * d)
    for name, sl in (("dq", slice(0, d)), ("dk", slice(d, 2 * d)), ("dv", slice(2 * d, 3 * d))):
        g, w = dqkv[:, sl], want[:, sl]
        err = (g - w).abs().max().item()
        assert err <= 0.02 * w.abs().max().item() + 2e-3, (name, err, w.abs().max().item())
        cos = torch.nn.functional.cosine_similarity(g.flatten(), w.flatten(), dim=0).item()
        assert cos > 0.999, (name, cos)


# ----------------------------------------------------------------------------- train-mode dropout (models_archs.py:51,58,135,187-199)
def _philox4x32_10(c, k):
    """NumPy Philox4x32-10 (Salmon et al., as curand / torch use it): c (n, 4) uint32 counters, k (2,) key."""
    c = c.astype(np.uint64).copy()
    k0, k1 = np.uint64(k[0]), np.uint64(k[1])
    M0, M1, W0, W1, MASK = np.uint64(0xD2511F53), np.uint64(0xCD9E8D57), np.uint64(0x9E3779B9), np.uint64(0xBB67AE85), np.uint64(0xFFFFFFFF)
    for _ in range(10):
        p0, p1 = M0 * c[:, 0], M1 * c[:, 2]
        hi0, lo0, hi1, lo1 = p0 >> np.uint64(32), p0 & MASK, p1 >> np.uint64(32), p1 & MASK
        c = np.stack([hi1 ^ c[:, 1] ^ k0, lo1, hi0 ^ c[:, 3] ^ k1, lo0], 1)
        k0, k1 = (k0 + W0) & MASK, (k1 + W1) & MASK
    return c.astype(np.uint32)


def _host_mask(rows, cols, seed, site, thr16):
    """The documented definition (include/vdr.h, vdr_dropout): element (r, c) kept iff the 16-bit lane c & 7 of
    Philox(counter = (c >> 3, r_lo, r_hi, site), key = seed) >= thr16."""
    r, c8 = np.meshgrid(np.arange(rows, dtype=np.uint64), np.arange((cols + 7) // 8, dtype=np.uint64), indexing="ij")
    ctr = np.stack([c8.ravel(), r.ravel() & np.uint64(0xFFFFFFFF), r.ravel() >> np.uint64(32), np.full(r.size, site, np.uint64)], 1)
    out = _philox4x32_10(ctr, (seed & 0xFFFFFFFF, seed >> 32)).reshape(rows, -1, 4)
    lanes = np.stack([out[..., 0] & 0xFFFF, out[..., 0] >> 16, out[..., 1] & 0xFFFF, out[..., 1] >> 16,
                      out[..., 2] & 0xFFFF, out[..., 2] >> 16, out[..., 3] & 0xFFFF, out[..., 3] >> 16], -1).reshape(rows, -1)[:, :cols]
    return lanes >= thr16


@pytest.mark.parametrize("p", [0.1, 0.5])
def test_dropout_mask_definition_and_statistics(cuda, p):
    """The device mask == the documented Philox definition (bit for bit), keep rate = 1 - p within binomial noise, kept values
    scaled by 1 / (1 - p), and p = 0 is the identity."""
    from vit_deep_radiomics_b200 import ops
    d = ops.Drop(seed=0x1234567890ABCDEF, site=7, p=p)
    m = ops.dropout_mask(300, 1000, d, cuda).cpu().numpy().astype(bool)
    assert np.array_equal(m, _host_mask(300, 1000, d.seed, d.site, d.thr16))
    keep = m.mean()
    assert abs(keep - (1 - d.p)) < 4 * np.sqrt(d.p * (1 - d.p) / m.size)
    assert not np.array_equal(m, ops.dropout_mask(300, 1000, ops.Drop(d.seed, 8, p), cuda).cpu().numpy().astype(bool))   # another site, another mask
    x = torch.randn(64, 256, device=cuda).bfloat16()
    y = ops.dropout_apply(x, d).float().cpu().numpy()
    mk = ops.dropout_mask(64, 256, d, cuda).cpu().numpy().astype(bool)
    want = np.where(mk, x.float().cpu().numpy() / (1 - d.p), 0.0)
    assert np.allclose(y, want, rtol=2 ** -7, atol=0) and np.all(y[~mk] == 0)
    z = torch.randn(32, 64, device=cuda).bfloat16()
    assert torch.equal(ops.gelu(z), ops.gelu(z, drop=ops.Drop(1, 1, 0.0)))                                                 # p = 0: bit-identical


def test_dropout_sites_match_exported_mask(cuda):
    """Every kernel that applies dropout uses the exported mask: GELU (fwd / bwd), the residual GEMM epilogue, the head."""
    from vit_deep_radiomics_b200 import ops
    torch.manual_seed(0)
    d = ops.Drop(seed=99, site=3, p=0.3)
    z = torch.randn(130, 256, device=cuda).bfloat16()
    mk = ops.dropout_mask(130, 256, d, cuda).bool()
    sc = 1.0 / (1 - d.p)
    h, h0 = ops.gelu(z, drop=d).float(), ops.gelu(z).float()
    assert torch.all(h[~mk] == 0) and torch.allclose(h[mk], h0[mk] * sc, rtol=2 ** -7, atol=1e-6)
    dh = torch.randn(130, 256, device=cuda).bfloat16()
    dz, dz0 = ops.gelu_bwd(dh, z, drop=d).float(), ops.gelu_bwd(dh, z).float()
    assert torch.all(dz[~mk] == 0) and torch.allclose(dz[mk], dz0[mk] * sc, rtol=2 ** -6, atol=1e-6)
    # C = R + dropout(A W^T + b)
    a = (torch.randn(130, 64, device=cuda) * 0.5).bfloat16()
    w = (torch.randn(256, 64, device=cuda) * 0.2).bfloat16()
    b = torch.randn(256, device=cuda)
    r = torch.randn(130, 256, device=cuda).bfloat16()
    got = ops.gemm(a, w, b, epilogue="residual", residual=r, drop=d).float()
    lin = a.float() @ w.float().t() + b
    want = r.float() + torch.where(mk, lin * sc, torch.zeros_like(lin))
    assert (got - want).abs().max() < 0.05
    plain = ops.gemm(a, w, b, epilogue="residual", residual=r).float()
    assert torch.equal(ops.gemm(a, w, b, epilogue="residual", residual=r, drop=ops.Drop(5, 5, 0.0)).float(), plain)        # p = 0 path untouched
    # head: hidden mask = row 0, logit mask = row 1 of the site
    cls = torch.randn(64, device=cuda).bfloat16()
    W1, b1, W2, b2 = torch.randn(128, 64, device=cuda) * 0.2, torch.randn(128, device=cuda) * 0.1, torch.randn(2, 128, device=cuda) * 0.2, torch.randn(2, device=cuda)
    hm = ops.dropout_mask(2, 128, d, cuda).bool()
    logits, zc = ops.cls_head_fwd(cls, W1, b1, W2, b2, drop=d)
    hid = torch.nn.functional.gelu(W1 @ cls.float() + b1) * hm[0].float() * sc
    want_l = (W2 @ hid + b2) * hm[1, :2].float() * sc
    assert torch.allclose(logits, want_l, rtol=1e-3, atol=1e-4)


@pytest.mark.parametrize("N,heads,p", [(200, 2, 0.1), (1025, 1, 0.5), (130, 4, 0.25)])
def test_flash_attention_dropout_fwd_bwd_vs_fp32_autograd(cuda, N, heads, p):
    """Attention dropout (nn.MultiheadAttention(dropout=p)): softmax -> mask / (1 - p) -> P V, forward AND the fused backward
    (which regenerates the mask) against fp32 autograd using the exported mask; lse is the dropout-free normaliser."""
    from vit_deep_radiomics_b200 import ops
    torch.manual_seed(N)
    d = heads * 64
    qkv = (torch.randn(N, 3 * d, device=cuda) * 0.7).bfloat16()
    drop = ops.Drop(seed=4242, site=11, p=p)
    out, lse = ops.flash_attn(qkv, 1, N, heads, return_lse=True, drop=drop)
    mk = ops.dropout_mask(heads * N, N, drop, cuda).bool().view(heads, N, N)          # row = (b*heads + h)*N + q
    x = qkv.float().clone().requires_grad_(True)
    q, k, v = (x[:, i * d:(i + 1) * d].view(N, heads, 64).transpose(0, 1) for i in range(3))
    s = q @ k.transpose(1, 2) / 8.0
    pm = torch.softmax(s, -1) * mk.float() / (1 - drop.p)
    ref = (pm @ v).transpose(0, 1).reshape(N, d)
    assert (out.float() - ref).abs().max() < 0.06
    assert torch.allclose(lse[0], torch.logsumexp(s, -1).detach(), rtol=0, atol=2e-2)
    do = torch.randn(N, d, device=cuda).bfloat16()
    ref.backward(do.float())
    dqkv = ops.flash_attn_bwd(qkv, out, do, lse, 1, N, heads, drop=drop).float()
    for i, name in enumerate("qkv"):
        g, w = dqkv[:, i * d:(i + 1) * d].flatten().double(), x.grad[:, i * d:(i + 1) * d].flatten().double()
        cos = float(g @ w / (g.norm() * w.norm()))
        assert cos > 0.998, (name, cos)
    # without dropout the same call is the p = 0 kernel
    assert torch.equal(ops.flash_attn(qkv, 1, N, heads), ops.flash_attn(qkv, 1, N, heads, drop=ops.Drop(1, 2, 0.0)))


@pytest.mark.parametrize("gain,nats", [(5.0, 35), (16.0, 120)])
def test_flash_attention_dropout_maximum_free_blocks_guard(cuda, gain, nats):
    """The dropout instantiation runs maximum-free key blocks as well: late keys tens of nats above the block-0 maximum (large P) and
    hundreds of nats above it (overflow -> the tile is recomputed exactly, dropout mask included) against fp32 with the exported mask."""
    from vit_deep_radiomics_b200 import ops
    torch.manual_seed(int(gain))
    N, heads = 700, 2
    d = heads * 64
    qkv = torch.randn(N, 3 * d, device=cuda)
    x = qkv.view(N, 3, heads, 64)
    x[:, 0] = x[:, 0].abs() * 0.5 + 2.0
    x[:, 1] = x[:, 1].abs() * 0.1 + 0.5
    x[200:520, 1] *= gain
    qkv = qkv.bfloat16()
    drop = ops.Drop(seed=99, site=3, p=0.25)
    out, lse = ops.flash_attn(qkv, 1, N, heads, return_lse=True, drop=drop)
    mk = ops.dropout_mask(heads * N, N, drop, cuda).bool().view(heads, N, N)
    q, k, v = (qkv.float()[:, i * d:(i + 1) * d].view(N, heads, 64).transpose(0, 1) for i in range(3))
    s = q @ k.transpose(1, 2) / 8.0
    assert (s[..., 128:].amax(-1) - s[..., :128].amax(-1)).max().item() > nats
    ref = ((torch.softmax(s, -1) * mk.float() / (1 - drop.p)) @ v).transpose(0, 1).reshape(N, d)
    assert torch.isfinite(out.float()).all()
    assert (out.float() - ref).abs().max() < 0.06 * max(1.0, float(ref.abs().max()) / 3.0)
    assert torch.allclose(lse[0], torch.logsumexp(s, -1), rtol=2e-4, atol=2e-2)


def test_classifier_train_mode_dropout_gradients_vs_fp32_autograd(cuda):
    """TransformerNoduleClassifier in train() with the reference's rates (0.1 / 0.1): the kernels' forward and EVERY parameter
    gradient against an fp32 autograd restatement of the same network that applies the exported masks at the same sites
    (attention probabilities, both sub-layer outputs, the feed-forward activation, head hidden units, logits);
    eval() stays dropout-free and deterministic, train() differs between passes."""
    from oracle import classifier_fp32 as C
    from vit_deep_radiomics_b200 import classifier_kernels as ck, ops
    from vit_deep_radiomics_b200.models_archs import TransformerNoduleClassifier
    d, ff, heads, layers, n = 128, 256, 2, 2, 150
    sd0 = C.init_state_dict(d, ff, 2, layers, seed=21)
    model = TransformerNoduleClassifier(d, ff, heads, 2, layers)
    model.load_state_dict(sd0)
    model = model.to(cuda)
    x = torch.randn(1, n, d, generator=torch.Generator().manual_seed(3)).to(cuda)
    model.eval()
    with torch.no_grad():
        e1, e2 = model(x)[0], model(x)[0]
    assert torch.equal(e1, e2)
    model.train()
    with torch.no_grad():
        t1, t2 = model(x)[0], model(x)[0]
    assert not torch.equal(t1, t2) and not torch.equal(t1, e1)
    # one pass with a known seed through the autograd Function
    cfg = ck.DropCfg(seed=777, p=0.1, p_head=0.1)
    params = model.param_list()
    logits, cls = ck.ClassifierFunction.apply(x[0], heads, layers, cfg, *params)
    (logits * torch.tensor([1.0, -2.0], device=cuda)).sum().backward()
    N = n + 1

    def mask(site_drop, rows, cols):
        return ops.dropout_mask(rows, cols, site_drop, cuda).float() / (1 - site_drop.p)

    sd = {k: v.detach().clone().to(cuda).requires_grad_(True) for k, v in sd0.items()}
    y = torch.nn.functional.layer_norm(torch.cat([sd["cls_token"][0], x[0]], 0), (d,), sd["norm.weight"], sd["norm.bias"], 1e-5)
    for l in range(layers):
        pre = f"transformer_encoder.layers.{l}."
        qkv = y @ sd[pre + "self_attn.in_proj_weight"].t() + sd[pre + "self_attn.in_proj_bias"]
        q, k, v = (qkv[:, i * d:(i + 1) * d].view(N, heads, 64).transpose(0, 1) for i in range(3))
        pm = torch.softmax(q @ k.transpose(1, 2) / 8.0, -1) * mask(cfg.site(l, 0), heads * N, N).view(heads, N, N)
        a = (pm @ v).transpose(0, 1).reshape(N, d)
        t = y + (a @ sd[pre + "self_attn.out_proj.weight"].t() + sd[pre + "self_attn.out_proj.bias"]) * mask(cfg.site(l, 1), N, d)
        y1 = torch.nn.functional.layer_norm(t, (d,), sd[pre + "norm1.weight"], sd[pre + "norm1.bias"], 1e-5)
        h = torch.nn.functional.gelu(y1 @ sd[pre + "linear1.weight"].t() + sd[pre + "linear1.bias"]) * mask(cfg.site(l, 2), N, ff)
        u = y1 + (h @ sd[pre + "linear2.weight"].t() + sd[pre + "linear2.bias"]) * mask(cfg.site(l, 3), N, d)
        y = torch.nn.functional.layer_norm(u, (d,), sd[pre + "norm2.weight"], sd[pre + "norm2.bias"], 1e-5)
    hm = mask(cfg.head(0), 2, 2 * d)
    hid = torch.nn.functional.gelu(sd["classifier.dense1.weight"] @ y[0] + sd["classifier.dense1.bias"]) * hm[0]
    want = (sd["classifier.dense2.weight"] @ hid + sd["classifier.dense2.bias"]) * hm[1, :2]
    (want * torch.tensor([1.0, -2.0], device=cuda)).sum().backward()
    assert (logits.detach() - want.detach()).abs().max() < 0.05, (logits, want)
    worst = 1.0
    for name, p_ in model.named_parameters():
        g, w = p_.grad.detach().double().flatten(), sd[name].grad.double().flatten()
        assert torch.isfinite(g).all(), name
        if w.norm() < 1e-7:
            assert g.norm() < 1e-4, name
            continue
        cos = float(g @ w / (g.norm() * w.norm()))
        worst = min(worst, cos)
        assert cos > 0.99 and float((g - w).norm() / w.norm()) < 0.15, (name, cos)
    assert worst > 0.99
