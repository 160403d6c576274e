"""-m gpu: the assembled hot path against the oracle.
Floating-point tolerances (bf16 activations / fp32 accumulation vs the fp32 oracle), as BASELINE.json's
north_star asks: per-descriptor max-abs and relative error, token cosine >= 0.999; gather indices bit-exact."""
import os

import numpy as np
import pytest
import torch

import parity

pytestmark = pytest.mark.gpu

# per config: |descriptor error| (final-LayerNorm outputs are O(1)), rms relative error, min token cosine -- the stated bounds
# live in tests/parity.py (about twice what a B200 run measured), the measured values go to gpurun_out/parity_measured.json


def _oracle_dense(model, imgs_nchw):
    from oracle import vit_fp32
    with torch.no_grad():
        return vit_fp32.vit_forward(model.state_dict_f32, vit_fp32.VIT_CONFIGS[model.model_name], imgs_nchw).numpy()


def _check_descriptors(got, want, key):
    return parity.check(key, got, want)


@pytest.mark.parametrize("name,hw,B", [("vit_t16", (64, 64), 3), ("vit_t16", (48, 80), 2), ("vit_s16", (224, 224), 8)])
def test_vit_dense_descriptors_vs_oracle(cuda, name, hw, B):
    """C1 (ViT-S/16, 224^2, batch 8) and tiny / non-square variants."""
    from vit_deep_radiomics_b200 import tfds_dense_descriptor as tdd
    torch.manual_seed(0)
    model = tdd.load_model(name, img_hw=hw, device=cuda, seed=11)
    x = torch.rand(B, 3, *hw)
    got = model.dense_descriptors(x.to(cuda)).cpu().numpy()
    want = _oracle_dense(model, x)
    assert got.shape == want.shape == (B, hw[0] // 16, hw[1] // 16, model.cfg["dim"])
    _check_descriptors(got, want, f"descriptors {name}@{hw[0]}x{hw[1]} B{B}")


def test_vit_l14_patch14_k_tail(cuda):
    """ViT-L/14 geometry (K = 588 not a multiple of 64, 24 layers) at a small image: 56x56 -> 4x4 patches."""
    from vit_deep_radiomics_b200 import tfds_dense_descriptor as tdd
    model = tdd.load_model("vit_l14", img_hw=(56, 56), device=cuda, seed=5)
    x = torch.rand(2, 3, 56, 56)
    _check_descriptors(model.dense_descriptors(x.to(cuda)).cpu().numpy(), _oracle_dense(model, x), "descriptors vit_l14@56x56 B2")


def test_get_dense_descriptor_single_slice_api(cuda):
    from vit_deep_radiomics_b200 import tfds_dense_descriptor as tdd
    model = tdd.load_model("vit_t16", img_hw=(64, 64), device=cuda, seed=2)
    img = np.random.default_rng(0).random((64, 64)).astype(np.float32)       # gray slice in 0..1
    f = tdd.get_dense_descriptor(model, img)
    assert f.shape == (4, 4, 128) and f.dtype == np.float32
    want = _oracle_dense(model, torch.from_numpy(np.stack([img] * 3))[None])
    _check_descriptors(f[None], want, "get_dense_descriptor vit_t16@64x64")


@pytest.mark.parametrize("case", ["T0", "C1"])
def test_extract_point_cloud_vs_oracle(cuda, case):
    """Fused extraction + gather == oracle generate_features + _get_features on the oracle's fp32 descriptors:
    identical token set / order / coordinates (bit-exact), descriptors within tolerance."""
    from oracle import gather_np as G
    from vit_deep_radiomics_b200 import synth, tfds_dense_descriptor as tdd
    img, mask, res, name = synth.make_case(case)
    H, W, S = img.shape
    model = tdd.load_model(name, img_hw=(H, W), device=cuda, seed=7)
    noise = (1.5, -2.0, 0.25)
    out = tdd.extract_point_cloud(model, img, mask, res, noise=noise)

    def oracle_descriptor(img2d):
        return _oracle_dense(model, torch.from_numpy(np.stack([img2d] * 3))[None].float())[0]

    feats, masks = G.generate_features(oracle_descriptor, img, mask)
    ref = G.token_gather(feats, masks, res, noise)
    assert out["count"] == ref["flat"].size > 0
    assert np.array_equal(out["src"].numpy(), ref["src"])
    _check_descriptors(out["tokens"].numpy()[None], ref["tokens"][None], f"point-cloud tokens {case}")
    # generate_features API: same shapes / masks as the oracle's, descriptors within tolerance
    fl, ml = tdd.generate_features(model, img, mask)
    assert len(fl) == len(feats) == S
    assert all(a.shape == b.shape for a, b in zip(fl, feats)) and all(np.array_equal(a, b) for a, b in zip(ml, masks))
    _check_descriptors(np.stack(fl), np.stack(feats), f"generate_features {case}")
    # gathering the device descriptors with the oracle gives the device tokens bit-for-bit (PE aside): the
    # gather itself adds no error
    ref2 = G.token_gather(fl, ml, res, noise)
    assert np.allclose(out["tokens"].numpy(), ref2["tokens"].astype(np.float32), rtol=0, atol=1e-6)


@pytest.mark.parametrize("case,n_slices", [("C2", 3), ("C4", 3)])
def test_headline_config_descriptors_and_point_cloud_vs_oracle(cuda, case, n_slices):
    """FLOAT parity at the benchmarked configurations themselves: ViT-B/16 @ 512x512 (C2, N = 1025 tokens per slice, the bench
    line's workload) and ViT-L/14 @ 224x224 (C4, 24 layers, K = 588 patch rows) -- the middle slices of the config's own synthetic
    volume through the batched device path vs oracle/vit_fp32.py in fp32 on the CPU, dense descriptors AND the mask-gathered
    point cloud (indices bit-exact, tokens within the stated bounds of tests/parity.py).
    reference: tfds_dense_descriptor.py:110-139 (get_dense_descriptor), train_models.py:143-182 (_get_features)."""
    from oracle import gather_np as G
    from vit_deep_radiomics_b200 import synth, tfds_dense_descriptor as tdd
    img, mask, res, name = synth.make_case(case)
    H, W, S = img.shape
    s0 = S // 2 - n_slices // 2
    img, mask = np.ascontiguousarray(img[:, :, s0:s0 + n_slices]), np.ascontiguousarray(mask[:, :, s0:s0 + n_slices])
    model = tdd.load_model(name, img_hw=(H, W), device=cuda, seed=1234)
    x = torch.from_numpy(np.ascontiguousarray(np.moveaxis(img, -1, 0)))[:, None].expand(-1, 3, -1, -1).contiguous()
    want = _oracle_dense(model, x)
    got = model.dense_descriptors(x.to(cuda)).cpu().numpy()
    p = model.cfg["patch"]
    assert got.shape == want.shape == (n_slices, H // p, W // p, model.cfg["dim"])
    _check_descriptors(got, want, f"descriptors {name}@{H}x{W} ({case})")
    # the fused path on the same slices: volume staging (TMA im2col view for C2) -> batched forward -> ROI -> gather
    out = tdd.extract_point_cloud(model, img, mask, res)
    order = iter(range(n_slices))

    def oracle_descriptor(img2d):             # the crop window of these configs is the whole slice: reuse the batch computed above
        k = next(order)
        assert np.array_equal(img2d, img[:, :, k])
        return want[k]

    feats, masks = G.generate_features(oracle_descriptor, img, mask)
    ref = G.token_gather(feats, masks, res)
    assert out["count"] == ref["flat"].size > 50
    assert np.array_equal(out["src"].numpy(), ref["src"])
    _check_descriptors(out["tokens"].numpy(), ref["tokens"], f"point-cloud tokens {name}@{H}x{W} ({case})")


def test_c2_full_size_properties(cuda):
    """BASELINE config C2 (ViT-B/16, 512x512x120): size-independent properties of the gather at full size."""
    from vit_deep_radiomics_b200 import ops, synth, tfds_dense_descriptor as tdd
    img, mask, res, name = synth.make_case("C2")
    H, W, S = img.shape
    model = tdd.load_model(name, img_hw=(H, W), device=cuda, seed=1234)
    out = tdd.extract_point_cloud(model, img, mask, res, add_pe=False, to_host=False)
    n = int(out["count"].item())
    src = out["src"][:n].cpu().numpy().astype(np.int64)
    fy0, fy1, fx0, fx1 = out["plan"]["feat_roi"]
    my0, my1, mx0, mx1 = out["plan"]["mask_roi"]
    h, w = fy1 - fy0, fx1 - fx0
    # expected selection from the host mask with the oracle-validated index maps
    rm, cm = ops.nearest_index_map(h, my1 - my0), ops.nearest_index_map(w, mx1 - mx0)
    sel = mask[my0:my1, mx0:mx1][np.ix_(rm, cm)]                    # (h, w, S)
    assert n == int(sel.sum()) and 3000 < n < 8000
    flat = src[:, 1] * (w * S) + src[:, 2] * S + src[:, 0]
    assert np.all(np.diff(flat) > 0)                                 # stable, ascending reference order
    assert sel.reshape(-1)[flat].all()
    # payload: gathered rows are exactly the rows of the final token matrix
    tok = model._ws[S]["OUT"].view(S, model.n_tokens, -1)
    rows = tok[torch.from_numpy(src[:, 0]).to(cuda), torch.from_numpy(1 + (src[:, 1] + fy0) * model.grid[1] + src[:, 2] + fx0).to(cuda)]
    assert torch.equal(out["tokens"][:n], rows)
    assert torch.isfinite(out["tokens"][:n]).all()


def test_classifier_forward_vs_golden(cuda, golden_dir):
    """Point-cloud classifier forward (kernels) against the reference's own outputs (golden) -- bf16 tolerance."""
    from vit_deep_radiomics_b200.models_archs import TransformerNoduleClassifier
    g = np.load(os.path.join(golden_dir, "classifier_small.npz"))
    d, ff, heads, layers = g["cfg"].tolist()
    model = TransformerNoduleClassifier(d, ff, heads, 2, layers)
    model.load_state_dict({k[len("param__"):]: torch.tensor(g[k]) for k in g.files if k.startswith("param__")})
    model = model.to(cuda).eval()
    with torch.no_grad():
        logits, cls = model(torch.tensor(g["x"]).to(cuda))
    assert logits.shape == (1, 2) and cls.shape == (1, d)
    parity.check("classifier logits (golden, small)", logits.cpu().numpy(), g["logits"], dict(abs=0.03, rel=1.0, cos=-1.0))
    parity.check("classifier CLS (golden, small)", cls.cpu().numpy(), g["cls"], dict(abs=0.06, rel=1.0, cos=0.999))


def test_classifier_full_config_vs_oracle(cuda):
    """Reference hyper-parameters (d 256, ff 1024, 4 heads, 2 layers) on a 2,000-token cloud vs the fp32 oracle."""
    from oracle import classifier_fp32 as C
    from vit_deep_radiomics_b200.models_archs import TransformerNoduleClassifier
    sd = C.init_state_dict(256, 1024, 2, 2, seed=4)
    model = TransformerNoduleClassifier(256, 1024, 4, 2, 2)
    model.load_state_dict(sd)
    model = model.to(cuda).eval()
    x = torch.randn(1, 2000, 256, generator=torch.Generator().manual_seed(1))
    with torch.no_grad():
        logits, cls = model(x.to(cuda))
        want_l, want_c = C.classifier_forward(sd, x, 4, 2)
    parity.check("classifier logits (d256 ff1024 h4 L2, n=2000)", logits.cpu().numpy(), want_l.numpy(), dict(abs=0.05, rel=1.0, cos=-1.0))
    parity.check("classifier CLS (d256 ff1024 h4 L2, n=2000)", cls.cpu().numpy(), want_c.numpy(), dict(abs=1.0, rel=1.0, cos=0.999))


def test_classifier_backward_vs_golden(cuda, golden_dir):
    """loss.backward() through the kernels against the reference's own gradients (golden): FocalLoss on the
    logits, every parameter gradient compared by cosine and relative error (bf16 operands, f32 accumulate)."""
    from vit_deep_radiomics_b200.models_archs import TransformerNoduleClassifier
    from vit_deep_radiomics_b200.train_models import FocalLoss
    g = np.load(os.path.join(golden_dir, "classifier_small.npz"))
    d, ff, heads, layers = g["cfg"].tolist()
    model = TransformerNoduleClassifier(d, ff, heads, 2, layers)
    model.load_state_dict({k[len("param__"):]: torch.tensor(g[k]) for k in g.files if k.startswith("param__")})
    from vit_deep_radiomics_b200.models_archs import set_dropout
    model = set_dropout(model.to(cuda), 0.0, 0.0).train()      # the golden gradients come from the reference in eval() mode (no dropout)
    crit = FocalLoss(alpha=torch.tensor([0.25, 0.75], device=cuda), gamma=2)
    logits, cls = model(torch.tensor(g["x"]).to(cuda))
    loss = crit(torch.squeeze(logits), torch.tensor(g["y"]).to(cuda)[0])
    assert abs(loss.item() - float(g["loss"])) < 0.02
    loss.backward()
    worst = 1.0
    for name, p in model.named_parameters():
        want = torch.tensor(g["grad__" + name]).double().flatten()
        got = p.grad.detach().cpu().double().flatten()
        assert torch.isfinite(got).all(), name
        if want.norm() < 1e-7:
            assert got.norm() < 1e-4, name
            continue
        cos = float((got @ want) / (got.norm() * want.norm()))
        rel = float((got - want).norm() / want.norm())
        worst = min(worst, cos)
        parity.check("classifier gradients (golden, small)", got.numpy()[None], want.numpy()[None], dict(abs=1e9, rel=0.12, cos=0.99))
    assert worst > 0.99


def test_train_step_accumulation_semantics(cuda):
    """train_epoch reproduces the reference loop (train_models.py:652-688): loss / iters, optimizer step every
    iters samples and at the last one -- checked against the fp32 oracle trained the same way on CPU."""
    from oracle import classifier_fp32 as C
    from vit_deep_radiomics_b200.models_archs import TransformerNoduleClassifier
    from vit_deep_radiomics_b200.train_models import FocalLoss, train_epoch
    torch.manual_seed(0)
    sd0 = C.init_state_dict(64, 128, 2, 1, seed=9)
    from vit_deep_radiomics_b200.models_archs import set_dropout
    model = TransformerNoduleClassifier(64, 128, 1, 2, 1)
    model.load_state_dict(sd0)
    model = set_dropout(model.to(cuda), 0.0, 0.0)              # parity with the dropout-free fp32 oracle
    opt = torch.optim.AdamW(model.parameters(), lr=5e-4, weight_decay=0.01)
    gen = torch.Generator().manual_seed(2)
    data = [(torch.randn(int(n), 64, generator=gen), torch.eye(2)[int(c)]) for n, c in [(40, 0), (25, 1), (33, 1), (17, 0), (29, 1)]]
    crit = FocalLoss(alpha=torch.tensor([0.25, 0.75], device=cuda), gamma=2)
    loss_gpu, _ = train_epoch(model, [(x.to(cuda), y.to(cuda)) for x, y in data], crit, opt, virtual_batch_size=2)
    # oracle: same loop in fp32 on CPU
    sd = {k: v.clone().requires_grad_(True) for k, v in sd0.items()}
    opt_c = torch.optim.AdamW(list(sd.values()), lr=5e-4, weight_decay=0.01)
    iters, tot = 2, 0.0
    for i, (x, y) in enumerate(data):
        lg, _ = C.classifier_forward(sd, x[None], 1, 1)
        l = C.focal_loss(lg[0], y, 2.0, torch.tensor([0.25, 0.75])) / iters
        l.backward()
        tot += l.item() * iters
        if (i + 1) % iters == 0 or i + 1 == len(data):
            opt_c.step()
            opt_c.zero_grad()
    assert abs(loss_gpu - tot / len(data)) < 0.02
    for k, v in model.state_dict().items():
        assert (v.cpu() - sd[k].detach()).abs().max() < 3e-3, k     # 3 AdamW steps of lr 5e-4: updates ~1.5e-3


def test_cuda_graph_training_step_equals_eager_step(cuda):
    """graph_step.GraphedTrainStep: three epochs of the accumulation loop with the per-length CUDA graphs (first visit eager, second
    visit captured + replayed, third visit replayed; weights change in between) give the weights of the eager loop (to the last bits) --
    without dropout, and with train-mode dropout when the eager model draws from the same (seed, counter) sequence."""
    from vit_deep_radiomics_b200 import classifier_kernels as ck
    from vit_deep_radiomics_b200.graph_step import graphed_step
    from vit_deep_radiomics_b200.models_archs import TransformerNoduleClassifier, set_dropout
    from vit_deep_radiomics_b200.train_models import FocalLoss, train_epoch
    gen = torch.Generator().manual_seed(5)
    data = [(torch.randn(int(n), 128, generator=gen).to(cuda), torch.eye(2)[int(c)].to(cuda)) for n, c in [(200, 0), (77, 1), (200, 1), (131, 0), (77, 0)]]
    crit = FocalLoss(alpha=torch.tensor([0.25, 0.75], device=cuda), gamma=2)
    for p_drop in (0.0, 0.1):
        out = []
        for graphs in (False, True):
            torch.manual_seed(1)
            model = set_dropout(TransformerNoduleClassifier(128, 256, 2, 2, 2).to(cuda), p_drop, p_drop)
            opt = torch.optim.AdamW(model.parameters(), lr=1e-3, weight_decay=0.01)
            if p_drop > 0:
                # one seed sequence for both runs: seed = base + (number of samples stepped so far), the graph's device counter
                step = graphed_step(model, crit)
                step.base_seed = 1234
                if not graphs:
                    k = [0]

                    def override(m=model, k=k):
                        k[0] += 1
                        return ck.DropCfg(1234 + k[0], *m._drop_rates())
                    model._drop_override = override
                else:
                    # eager visits inside the graphed loop must consume the counter as the captured ones do
                    def eager(x, label, inv, s=step, m=model):
                        s.seed_offset.add_(1)
                        m._drop_override = lambda: ck.DropCfg(s.base_seed, *m._drop_rates(), seed_offset=s.seed_offset)
                        try:
                            return type(s)._eager(s, x, label, inv)
                        finally:
                            m._drop_override = None
                    step._eager = eager
            losses = [train_epoch(model, data, crit, opt, virtual_batch_size=2, cuda_graphs=graphs)[0] for _ in range(3)]
            if graphs:
                st = graphed_step(model, crit)
                assert st.replays >= 9 and len(st.graphs) == 3, (st.replays, st.eager, len(st.graphs))
            out.append((losses, {k_: v.clone() for k_, v in model.state_dict().items()}))
        (l0, w0), (l1, w1) = out
        # same kernels in the same order; the backward's float atomics (dq, LayerNorm column sums) make two runs differ in the last bits
        assert max(abs(a - b) for a, b in zip(l0, l1)) < 1e-5, (p_drop, l0, l1)
        for k_ in w0:
            assert (w0[k_] - w1[k_]).abs().max() < 2e-5, (p_drop, k_, float((w0[k_] - w1[k_]).abs().max()))


def test_train_epoch_bimodal_crossmodal_loss_matches_cpu_training(cuda, golden_dir):
    """The reference's bimodal loop (train_models.py:656-688 with loss_func 'crossmodal'): model(ct, pet) -> CrossModalFocalLoss on
    outputs[0], [2], [3] / iters, optimizer steps as in the unimodal loop -- against the fp32 oracle trained the same way on CPU."""
    from oracle import classifier_fp32 as C
    from vit_deep_radiomics_b200.models_archs import TransformerNoduleBimodalClassifier
    from vit_deep_radiomics_b200.train_models import CrossModalFocalLoss, make_criterion, train_epoch
    g = np.load(os.path.join(golden_dir, "bimodal_small.npz"))
    cfg = [int(v) for v in g["cfg"]]
    d, heads_ct, heads_pet, layers_ct, layers_pet = cfg[0], cfg[3], cfg[4], cfg[5], cfg[6]
    sd0 = {k[len("param__"):]: torch.tensor(g[k]) for k in g.files if k.startswith("param__")}
    from vit_deep_radiomics_b200.models_archs import set_dropout
    model = TransformerNoduleBimodalClassifier(*cfg)
    model.load_state_dict(sd0)
    model = set_dropout(model.to(cuda), 0.0, 0.0)              # parity with the dropout-free fp32 oracle
    opt = torch.optim.AdamW(model.parameters(), lr=5e-4, weight_decay=0.01)
    gen = torch.Generator().manual_seed(4)
    data = [(torch.randn(nc, d, generator=gen), torch.randn(np_, d, generator=gen), torch.eye(2)[c])
            for nc, np_, c in [(30, 12, 0), (21, 40, 1), (17, 9, 1), (26, 26, 0)]]
    crit = make_criterion("crossmodal", cuda)
    loss_gpu, scores = train_epoch(model, [tuple(t.to(cuda) for t in s_) for s_ in data], crit, opt, virtual_batch_size=3)
    assert len(scores) == 4 and all(abs(float(sc.sum()) - 1) < 1e-5 for sc in scores)
    sd = {k: v.clone().requires_grad_(True) for k, v in sd0.items()}
    opt_c = torch.optim.AdamW(list(sd.values()), lr=5e-4, weight_decay=0.01)
    crit_c = CrossModalFocalLoss(alpha=torch.tensor([0.25, 0.75]), gamma_unimodal=2.0, gamma_bimodal=1.0, beta=0.6)
    iters, tot = 3, 0.0
    for i, (xc, xp, y) in enumerate(data):
        lg, _, lc, lp = C.bimodal_forward(sd, xc[None], xp[None], heads_ct, heads_pet, layers_ct, layers_pet)
        l = crit_c(lg[0], lc[0], lp[0], y) / iters
        l.backward()
        tot += l.item() * iters
        if (i + 1) % iters == 0 or i + 1 == len(data):
            opt_c.step()
            opt_c.zero_grad()
    assert abs(loss_gpu - tot / len(data)) < 0.02
    for k, v in model.state_dict().items():
        assert (v.cpu() - sd[k].detach()).abs().max() < 3e-3, k     # 2 AdamW steps of lr 5e-4


def test_cuda_graph_training_step_bimodal_equals_eager_step(cuda, golden_dir):
    """graph_step.GraphedTrainStep for the bimodal model: one graph per (CT count, PET count) pair, CrossModalFocalLoss on
    outputs[0], [2], [3]; three epochs (eager visit, capture, replay -- weights change in between) give the eager loop's losses and
    weights (float atomics in the backward: last bits)."""
    from vit_deep_radiomics_b200.graph_step import graphed_step
    from vit_deep_radiomics_b200.models_archs import TransformerNoduleBimodalClassifier, set_dropout
    from vit_deep_radiomics_b200.train_models import make_criterion, train_epoch
    g = np.load(os.path.join(golden_dir, "bimodal_small.npz"))
    cfg = [int(v) for v in g["cfg"]]
    d = cfg[0]
    sd0 = {k[len("param__"):]: torch.tensor(g[k]) for k in g.files if k.startswith("param__")}
    gen = torch.Generator().manual_seed(9)
    data = [(torch.randn(nc, d, generator=gen).to(cuda), torch.randn(np_, d, generator=gen).to(cuda), torch.eye(2)[c].to(cuda))
            for nc, np_, c in [(30, 12, 0), (21, 40, 1), (30, 12, 1), (26, 26, 0)]]
    crit = make_criterion("crossmodal", cuda)
    out = []
    for graphs in (False, True):
        model = TransformerNoduleBimodalClassifier(*cfg)
        model.load_state_dict(sd0)
        model = set_dropout(model.to(cuda), 0.0, 0.0)
        # (plain SGD: Adam's normalisation turns the float-atomics noise of mathematically-zero gradients -- key biases -- into lr-sized steps)
        opt = torch.optim.SGD(model.parameters(), lr=0.01)     # (the eager loop itself repeats to ~1e-5 in the loss at lr 0.05: atomics noise grows with the step size)
        losses = [train_epoch(model, data, crit, opt, virtual_batch_size=3, cuda_graphs=graphs)[0] for _ in range(3)]
        if graphs:
            st = graphed_step(model, crit)
            assert len(st.graphs) == 3 and st.replays >= 7, (st.replays, st.eager, len(st.graphs))   # (30, 12) is seen twice per epoch
        out.append((losses, {k_: v.clone() for k_, v in model.state_dict().items()}))
    (l0, w0), (l1, w1) = out
    assert max(abs(a - b) for a, b in zip(l0, l1)) < 1e-5, (l0, l1)
    for k_ in w0:
        assert (w0[k_] - w1[k_]).abs().max().item() < 2e-5, k_
    # train-mode dropout inside a captured bimodal step: replays draw fresh masks (device seed counter), losses stay finite
    model = TransformerNoduleBimodalClassifier(*cfg)
    model.load_state_dict(sd0)
    model = model.to(cuda)
    opt = torch.optim.AdamW(model.parameters(), lr=5e-4, weight_decay=0.01)
    ls = [train_epoch(model, data, crit, opt, virtual_batch_size=3, cuda_graphs=True)[0] for _ in range(4)]
    assert all(np.isfinite(v) for v in ls) and len(set(round(v, 6) for v in ls)) > 1


def test_native_vit_forward_equals_op_by_op_path(cuda):
    """vdr_vit_forward (one C call) enqueues the same kernels in the same order as the Python op-by-op path: identical tokens."""
    from vit_deep_radiomics_b200 import synth, tfds_dense_descriptor as tdd
    rng = np.random.default_rng(21)
    for name, hw, S in (("vit_t16", (256, 256), 5), ("vit_t16", (64, 64), 3)):     # TMA im2col view / materialised im2col
        model = tdd.load_model(name, img_hw=hw, device=cuda, seed=5)
        vol = torch.from_numpy(rng.random(hw + (S,), dtype=np.float32)).to(cuda)
        crop = (0, hw[0], 0, hw[1])
        model.use_native_forward = True
        a = model.forward_volume(vol, crop).clone()
        model.use_native_forward = False
        b = model.forward_volume(vol, crop).clone()
        model.use_native_forward = True
        assert torch.equal(a, b)
        assert torch.isfinite(a).all()


def test_folded_layernorm_forward_matches_layernorm_kernel_path(cuda):
    """LayerNorms folded into the qkv / fc1 GEMMs (default) against the path that runs the LayerNorm kernels: same tokens within
    bf16 noise, both within the oracle tolerance (test_vit_dense_descriptors_vs_oracle covers the default path), and run-to-run
    bit-identical (the statistics tables are written without atomics)."""
    from vit_deep_radiomics_b200 import tfds_dense_descriptor as tdd
    from vit_deep_radiomics_b200.vit import ViTBackbone
    rng = np.random.default_rng(3)
    for name, hw, S in (("vit_s16", (256, 256), 6), ("vit_t16", (64, 64), 3)):
        folded = tdd.load_model(name, img_hw=hw, device=cuda, seed=8)
        assert folded.fold_layernorm and "qkv_wf" in folded.w["blocks"][0]
        try:
            ViTBackbone.fold_layernorm = False
            plain = tdd.load_model(name, img_hw=hw, device=cuda, seed=8)
        finally:
            ViTBackbone.fold_layernorm = True
        assert "qkv_wf" not in plain.w["blocks"][0]
        vol = torch.from_numpy(rng.random(hw + (S,), dtype=np.float32)).to(cuda)
        crop = (0, hw[0], 0, hw[1])
        a = folded.forward_volume(vol, crop).clone()
        a2 = folded.forward_volume(vol, crop).clone()
        b = plain.forward_volume(vol, crop).clone()
        assert torch.equal(a, a2)
        _check_descriptors(a.cpu().numpy(), b.cpu().numpy(), f"folded vs unfolded LayerNorm {name}")
        x = torch.from_numpy(rng.random((2, 3) + hw, dtype=np.float32))
        _check_descriptors(plain.dense_descriptors(x.to(cuda)).cpu().numpy(), _oracle_dense(plain, x), f"descriptors unfolded {name}@{hw[0]}")


@pytest.mark.parametrize("ch,cw,oh,ow", [(37, 53, 64, 64), (200, 180, 256, 256), (96, 80, 64, 48), (300, 260, 128, 224)])
def test_resize_staging_matches_oracle(cuda, ch, cw, oh, ow):
    """vdr_volume_to_slices_resized == prepare_image's skimage resize (oracle/resize_np.py, pinned to scipy.ndimage):
    f32 arithmetic on the device, bf16 slices out -> within one bf16 rounding of the float64 oracle."""
    from oracle import resize_np
    from vit_deep_radiomics_b200 import ops
    rng = np.random.default_rng(ch + ow)
    H, W, S, y0, x0 = ch + 11, cw + 7, 5, 6, 3
    vol = rng.random((H, W, S), dtype=np.float32)
    got = ops.volume_to_slices(torch.from_numpy(vol).to(cuda), (y0, y0 + ch, x0, x0 + cw), out_hw=(oh, ow)).float().cpu().numpy()
    assert got.shape == (S, oh, ow)
    for s in range(S):
        want = resize_np.resize(vol[y0:y0 + ch, x0:x0 + cw, s], (oh, ow))
        assert np.abs(got[s] - want).max() <= 2 ** -8 * max(1.0, np.abs(want).max()) + 1e-4      # bf16 rounding + f32 arithmetic


def test_generate_features_resizes_crop_to_backbone_input(cuda):
    """A crop window of another size than the backbone input (the normal case on real data: the window is 4x the tumour
    bounding box) is resized on the device; descriptors match the fp32 oracle ViT run on the oracle-resized slices, and the
    ROI geometry matches the reference's extract_roi scaling."""
    from oracle import gather_np, resize_np
    from vit_deep_radiomics_b200 import tfds_dense_descriptor as tdd
    from vit_deep_radiomics_b200.visualization_utils import crop_window
    rng = np.random.default_rng(31)
    H = W = 160
    S = 4
    img = rng.random((H, W, S), dtype=np.float32)
    mask = np.zeros((H, W, S), dtype=bool)
    mask[70:85, 60:80, 1:3] = True                      # bbox 15 x 20 -> window side 80, backbone input 64
    model = tdd.load_model("vit_t16", img_hw=(64, 64), device=cuda, seed=3)
    feats, masks = tdd.generate_features(model, img, mask)
    x0, y0, x1, y1 = crop_window(mask.any(-1))
    y0, y1, x0, x1 = max(0, y0), min(H, y1), max(0, x0), min(W, x1)
    assert (y1 - y0, x1 - x0) != (64, 64)
    crop = img[y0:y1, x0:x1]
    resized = np.stack([resize_np.resize(crop[:, :, s], (64, 64)) for s in range(S)]).astype(np.float32)
    dense = _oracle_dense(model, torch.from_numpy(resized)[:, None].expand(-1, 3, -1, -1).contiguous())
    from vit_deep_radiomics_b200.visualization_utils import roi_window
    bigger_c = mask.any(-1)[y0:y1, x0:x1]
    fx0, fy0, fx1, fy1 = roi_window(model.grid, bigger_c, margin=1)
    mx0, my0, mx1, my1 = roi_window(bigger_c.shape, bigger_c, margin=1)
    assert len(feats) == S
    for s in range(S):
        assert feats[s].shape == dense[s, fy0:fy1, fx0:fx1].shape
        _check_descriptors(feats[s], dense[s, fy0:fy1, fx0:fx1], "generate_features resized crop")
        assert np.array_equal(masks[s], mask[y0:y1, x0:x1, s][my0:my1, mx0:mx1])


def test_c5_extraction_feeds_classifier_training_step(cuda):
    """BASELINE config C5 as a chain: device extraction -> point cloud (tokens stay on the GPU) -> classifier forward, focal
    loss, backward.  The classifier's logits / CLS token / gradients on the extracted tokens match the fp32 oracle classifier
    evaluated on the same tokens (tolerances of the classifier tests)."""
    from oracle import classifier_fp32 as C
    from vit_deep_radiomics_b200 import synth, tfds_dense_descriptor as tdd
    from vit_deep_radiomics_b200.models_archs import TransformerNoduleClassifier
    from vit_deep_radiomics_b200.train_models import FocalLoss
    img, mask, res, name = synth.make_case("T0")
    model = tdd.load_model(name, img_hw=img.shape[:2], device=cuda, seed=7)
    out = tdd.extract_point_cloud(model, img, mask, res, to_host=False)
    n = int(out["count"].item())
    assert n > 0
    tokens = out["tokens"][:n]
    D = model.cfg["dim"]
    torch.manual_seed(1)
    from vit_deep_radiomics_b200.models_archs import set_dropout
    clf = set_dropout(TransformerNoduleClassifier(D, 4 * D, D // 64, 2, 2).to(cuda), 0.0, 0.0)   # compared with the dropout-free oracle below
    y = torch.tensor([0.0, 1.0], device=cuda)
    logits, cls = clf(tokens.unsqueeze(0))
    loss = FocalLoss(alpha=torch.tensor([0.25, 0.75], device=cuda), gamma=2)(torch.squeeze(logits), y)
    loss.backward()
    sd = {k: v.detach().cpu().clone().requires_grad_(True) for k, v in clf.state_dict().items()}
    lg, cl = C.classifier_forward(sd, tokens.detach().cpu()[None], D // 64, 2)
    ref_loss = C.focal_loss(lg[0], y.cpu(), 2.0, torch.tensor([0.25, 0.75]))
    ref_loss.backward()
    assert torch.allclose(logits.detach().cpu().reshape(-1), lg.detach().reshape(-1), atol=0.05, rtol=0.05)
    assert torch.nn.functional.cosine_similarity(cls.detach().cpu().reshape(1, -1), cl.detach().reshape(1, -1)).item() > 0.999
    assert abs(float(loss) - float(ref_loss)) < 0.05 * max(1.0, abs(float(ref_loss)))
    for k, p in clf.named_parameters():
        assert p.grad is not None and torch.isfinite(p.grad).all(), k
    gw = dict(clf.named_parameters())["classifier.dense2.weight"].grad.cpu()
    assert torch.nn.functional.cosine_similarity(gw.reshape(1, -1), sd["classifier.dense2.weight"].grad.reshape(1, -1)).item() > 0.99


@pytest.mark.parametrize("mode", ["both", "ct", "pet"])
def test_bimodal_classifier_vs_golden(cuda, golden_dir, mode):
    """TransformerNoduleBimodalClassifier (scope row N4) through the kernels against the reference's own outputs and gradients
    (tests/golden/bimodal_small.npz, frozen from the unmodified module): both modalities (cross attention) and each one alone."""
    from vit_deep_radiomics_b200.models_archs import TransformerNoduleBimodalClassifier
    from vit_deep_radiomics_b200.train_models import FocalLoss
    g = np.load(os.path.join(golden_dir, "bimodal_small.npz"))
    from vit_deep_radiomics_b200.models_archs import set_dropout
    model = TransformerNoduleBimodalClassifier(*[int(v) for v in g["cfg"]])
    model.load_state_dict({k[len("param__"):]: torch.tensor(g[k]) for k in g.files if k.startswith("param__")})
    model = set_dropout(model.to(cuda), 0.0, 0.0).train()      # the golden gradients come from the reference without dropout
    x_ct = torch.tensor(g["x_ct"]).to(cuda) if mode in ("both", "ct") else None
    x_pet = torch.tensor(g["x_pet"]).to(cuda) if mode in ("both", "pet") else None
    y = torch.tensor(g["y"]).to(cuda)
    crit = FocalLoss(alpha=torch.tensor([0.25, 0.75], device=cuda), gamma=2)
    lg, z, lg_ct, lg_pet = model(x_ct, x_pet)
    assert lg.shape == (1, 2) and z.shape == (1, int(g["cfg"][0]))
    for got, key in ((lg, "logits"), (lg_ct, "logits_ct"), (lg_pet, "logits_pet")):
        assert np.abs(got.detach().cpu().numpy() - g[f"{mode}__{key}"]).max() < 0.04, key
    zc, zw = z.detach().cpu().numpy()[0].astype(np.float64), g[f"{mode}__z"].reshape(-1).astype(np.float64)
    assert (zc * zw).sum() / (np.linalg.norm(zc) * np.linalg.norm(zw)) > 0.999
    loss = crit(torch.squeeze(lg), y) + crit(torch.squeeze(lg_ct), y) + crit(torch.squeeze(lg_pet), y) + 0.1 * z.sum()
    assert abs(loss.item() - float(g[f"{mode}__loss"])) < 0.05
    loss.backward()
    checked = 0
    for name, p in model.named_parameters():
        key = f"{mode}__grad__{name}"
        if key not in g.files:
            assert p.grad is None or float(p.grad.abs().max()) == 0.0, name      # parameter not on this mode's path
            continue
        want = torch.tensor(g[key]).double().flatten()
        assert p.grad is not None, name
        got = p.grad.detach().cpu().double().flatten()
        assert torch.isfinite(got).all(), name
        if want.norm() < 1e-7:
            assert got.norm() < 1e-4, name
            continue
        cos = float((got @ want) / (got.norm() * want.norm()))
        rel = float((got - want).norm() / want.norm())
        parity.check(f"bimodal gradients (golden, {mode})", got.numpy()[None], want.numpy()[None], dict(abs=1e9, rel=0.15, cos=0.99))
        checked += 1
    assert checked >= 15


@pytest.mark.parametrize("augment", [False, True])
def test_dataset_items_on_device_match_oracle_gather(cuda, augment):
    """PETCTDataset3D (scope row N3) with the device gather: every item's CT / PET token sequences are bit-identical to the
    same dataset run with the CPU oracle gather (which tests/test_oracle_vs_reference.py pins to the reference's items)."""
    from oracle import gather_np as G
    from oracle import ref_shim
    from vit_deep_radiomics_b200 import train_models as tm
    D = 12
    df = tm.prepare_df(ref_shim.make_dataset_table(seed=5, D=D))
    enc = tm.get_label_encoder(df)
    kw = dict(use_augmentation=augment, feature_dim=D, arch="transformer", store=ref_shim.H5_FILES)
    dev_ds = tm.PETCTDataset3D(df, enc, "ct.h5", "pet.h5", device=cuda, **kw)
    cpu_ds = tm.PETCTDataset3D(df, enc, "ct.h5", "pet.h5", gather=lambda f, m, r, n, d: G.token_gather(f, m, r, n, d)["tokens"], **kw)
    assert len(dev_ds) == len(cpu_ds) > 0
    for i in range(len(dev_ds)):
        np.random.seed(7 + i)
        a = dev_ds[i]
        np.random.seed(7 + i)
        b = cpu_ds[i]
        assert a[0].is_cuda and a[1].is_cuda and a[3] == b[3] and torch.equal(a[2], b[2])
        assert torch.equal(a[0].cpu(), b[0]) and torch.equal(a[1].cpu(), b[1])


def test_run_fold_trains_the_shipped_classifier_config(cuda, tmp_path):
    """One fold of the training script (train_models.py:562-810) end to end on the device: dataset items gathered on the GPU,
    the classifier of conf/parameters_models.yaml (feature_dim 256) trained for three epochs with the reference's accumulation
    rule, evaluation pass, metric files, checkpoints.  The training loss of the fixed seed goes down."""
    import json
    from oracle import ref_shim
    from vit_deep_radiomics_b200 import config_manager, train_models as tm
    D = 256
    df = tm.prepare_df(ref_shim.make_dataset_table(seed=11, D=D))
    enc = tm.get_label_encoder(df)
    cfg = config_manager.load_conf(project_dir=os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
    assert cfg["models"]["transformer"]["feature_dim"] == D
    cfg["models"]["transformer"].update(patience=10, virtual_batch_size=4)
    torch.manual_seed(3)
    np.random.seed(3)
    hist = tm.run_fold(cfg, "transformer", "ct", df[df.patient_id.isin(["P1", "P2"])].reset_index(drop=True),
                       df[df.patient_id.isin(["P3", "P4"])].reset_index(drop=True), enc, "ct.h5", "pet.h5", str(tmp_path), kfold=2,
                       device=str(cuda), store=ref_shim.H5_FILES, num_epochs=3)
    assert [h["epoch"] for h in hist] == [0, 1, 2]
    assert all(np.isfinite(h["train_loss"]) and np.isfinite(h["test_loss"]) for h in hist)
    assert hist[-1]["train_loss"] < hist[0]["train_loss"]
    for e in range(3):
        for split in ("train", "test"):
            rep = json.load(open(tmp_path / f"{split}_metrics_{e}.json"))
            assert rep["split"] == split and rep["epoch"] == e and rep["kfold"] == 2 and 0.0 <= rep["ROC AUC"] <= 1.0
    saved = sorted(p.name for p in tmp_path.glob("model_epoch_*.pth"))
    assert saved and saved[0] == "model_epoch_0000.pth"          # epoch 0: target == the fold's mean
    sd = torch.load(tmp_path / saved[-1], map_location="cpu")
    assert "cls_token" in sd and sd["cls_token"].shape[-1] == D


def test_run_table_batches_small_patients_bit_identically(cuda):
    """PointCloudExtractor.run_table: several small patients share one backbone batch (forward_volumes); the table must be bit for
    bit what the one-patient-per-forward path writes (same kernels per row: the GEMM tiles never mix rows of different images)."""
    from vit_deep_radiomics_b200 import synth, tfds_dense_descriptor as tdd
    from vit_deep_radiomics_b200.distributed import PointCloudTable
    vols = [synth.make_case("C1", seed=50 + i) for i in range(3)]
    img, mask, res, name = vols[0]
    model = tdd.load_model(name, img_hw=img.shape[:2], device=cuda, seed=3)
    ex = tdd.PointCloudExtractor(model)
    items = [(i, torch.as_tensor(v[0]).to(cuda), torch.as_tensor(np.ascontiguousarray(v[1]).view(np.uint8)).to(cuda), v[2]) for i, v in enumerate(vols)]
    cap = 3 * img.shape[2] * model.grid[0] * model.grid[1]
    tables = []
    for max_rows in (0, 131072):
        tb = PointCloudTable(3, model.cfg["dim"], cap_rows=cap, device=cuda, rank=0, world=1)
        total = ex.run_table(items, tb, max_rows_per_batch=max_rows)
        tables.append((total, tb.tokens[:total].clone(), tb.src[:total].clone()))
    assert tables[0][0] == tables[1][0] and tables[0][0] > 30
    assert torch.equal(tables[0][2], tables[1][2])
    assert torch.equal(tables[0][1], tables[1][1])


@pytest.mark.gpu
@pytest.mark.parametrize("crop,out_hw", [((3, 59, 5, 61), (56, 56)), ((0, 70, 2, 100), (56, 84))])
def test_cell_padded_staging_layout(cuda, crop, out_hw):
    """vdr_volume_to_slices_cells: the pixels of the plain staging (same crop / resize), each 14 x 14 patch in the corner of a 16 x 16
    cell; every other pixel of the zeroed buffer stays zero."""
    from vit_deep_radiomics_b200 import ops
    g = torch.Generator().manual_seed(11)
    vol = torch.rand(72, 104, 5, generator=g).to(cuda)
    plain = ops.volume_to_slices(vol, crop, out_hw=out_hw)
    cells = ops.volume_to_slices(vol, crop, out_hw=out_hw, cells=(14, 16))
    S, OH, OW = plain.shape
    assert tuple(cells.shape) == (S, OH // 14 * 16, OW // 14 * 16)
    c = cells.view(S, OH // 14, 16, OW // 14, 16)
    assert torch.equal(c[:, :, :14, :, :14].reshape(S, OH, OW), plain)
    assert float(c[:, :, 14:, :, :].float().abs().max()) == 0.0 and float(c[:, :, :, :, 14:].float().abs().max()) == 0.0


@pytest.mark.gpu
def test_vit_l14_cell_padded_tma_patch_embed(cuda, monkeypatch):
    """ViT-L/14 at 224 x 224 (config C4): the cell-padded TMA im2col patch embedding (14-pixel patches in 16-pixel cells, zero weight
    columns at the pads) against the materialised-im2col path of the same weights -- same bf16 products, another accumulation order."""
    from vit_deep_radiomics_b200 import tfds_dense_descriptor as tdd
    g = torch.Generator().manual_seed(3)
    vol = torch.rand(230, 240, 3, generator=g).to(cuda)
    crop = (2, 226, 9, 233)
    model = tdd.load_model("vit_l14", img_hw=(224, 224), device=cuda, seed=5)
    assert model.cell == (14, 16) and model.stage_hw == (256, 256) and model.pe_patch == 16
    a = model.forward_volume(vol, crop).clone()
    model.use_native_forward = False                      # the op-by-op path reads the same cell-padded slices
    b = model.forward_volume(vol, crop).clone()
    monkeypatch.setenv("VDR_NO_CELL_PAD", "1")
    ref_model = tdd.load_model("vit_l14", img_hw=(224, 224), device=cuda, seed=5)
    assert ref_model.cell is None
    r = ref_model.forward_volume(vol, crop)
    assert torch.equal(a, b)
    # the patch embedding itself: same bf16 products, fp32 accumulation in another order -> at most a bf16 ulp of the embedded tokens
    from vit_deep_radiomics_b200 import ops
    S, N, d = 3, model.n_tokens, model.cfg["dim"]
    sl_cell = ops.volume_to_slices(vol, crop, out_hw=(224, 224), cells=(14, 16))
    sl = ops.volume_to_slices(vol, crop, out_hw=(224, 224))
    x_cell = torch.zeros(S * N, d, dtype=torch.bfloat16, device=cuda)
    x_ref = torch.zeros_like(x_cell)
    ops.patch_embed(sl_cell, model.w["pe_w_tma"], model.w["pe_b"], model.w["pos"], 16, out=x_cell)
    A = ops.im2col_gray_bf16(sl, 14)
    ops.gemm(A, ref_model.w["pe_w"], ref_model.w["pe_b"], epilogue="residual", residual=ref_model.w["pos"], out=x_ref, k=ref_model.K,
             out_group=(N - 1, N, 1), res_mod=(N - 1, 1))
    dx = float((x_cell.float() - x_ref.float()).abs().max())
    print("cell-padded TMA patch embedding vs materialised im2col GEMM: max abs %.3g (tokens of magnitude %.2f)" % (dx, float(x_ref.float().abs().max())))
    assert dx <= 2.0 ** -7 * max(1.0, float(x_ref.float().abs().max())), dx
    # ... and after 24 layers the two paths differ by what that last-bit difference grows to (the size of the bf16-vs-fp32 parity error)
    err = float((a - r).abs().max()), float((a - r).norm() / r.norm())
    print("cell-padded vs materialised path, descriptors (vit_l14@224): max abs %.3g, rms rel %.3g" % err)
    assert err[0] < parity.BOUNDS["descriptors vit_l14@224x224 (C4)"]["abs"] and err[1] < parity.BOUNDS["descriptors vit_l14@224x224 (C4)"]["rel"], err
