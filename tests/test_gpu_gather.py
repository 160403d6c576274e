"""-m gpu: the mask gathers through the C ABI -- bit-exact against the golden vectors frozen from the
reference and against the NumPy oracle on seeded inputs, incl. empty / full / ragged cases."""
import os

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu
CASES = ["square", "ragged", "wide", "probe", "empty", "full", "single"]


@pytest.mark.parametrize("name", CASES)
def test_g1_golden(cuda, golden_dir, name):
    from oracle import gather_np as G
    from vit_deep_radiomics_b200 import ops
    g = np.load(os.path.join(golden_dir, "gather_g1.npz"))
    feats, masks = g[f"{name}__features"], g[f"{name}__masks"]
    res, noise = g[f"{name}__res"], g[f"{name}__noise"]
    want = g[f"{name}__out_transformer"]                       # reference output (float64)
    f_t, m_t = torch.from_numpy(feats).to(cuda), torch.from_numpy(masks.astype(np.uint8)).to(cuda)
    tok, src, cnt = ops.mask_gather(f_t, m_t, pe=dict(res=res, noise=noise))
    n = int(cnt.item())
    assert n == want.shape[0]
    o = G.token_gather(list(feats), list(masks), res, noise)
    assert np.array_equal(src[:n].cpu().numpy(), o["src"])     # integer contract: bit-exact
    raw, _, _ = ops.mask_gather(f_t, m_t)
    assert np.array_equal(raw[:n].cpu().numpy(), o["raw"])     # gathered payload: bit-exact
    # tokens + PE/4: reference casts its float64 result to float32 (train_models.py:120); the device
    # computes the PE in fp64 with CUDA's sin/cos (<= 1 ulp f64) -> identical after the f32 cast up to 1 ulp f32
    got = tok[:n].cpu().numpy()
    assert np.allclose(got, want.astype(np.float32), rtol=0, atol=2.5e-7 * max(1.0, np.abs(want).max()) if n else 0)


def test_g1_roi_and_token_matrix_layout(cuda):
    """Gather straight out of a CLS-first token matrix with an ROI window == extract_roi + _get_features."""
    from oracle import gather_np as G
    from vit_deep_radiomics_b200 import ops
    rng = np.random.default_rng(8)
    S, gh, gw, D, HM, WM = 6, 8, 8, 64, 128, 128
    tok = rng.standard_normal((S, gh * gw + 1, D)).astype(np.float32)
    mask = rng.random((S, HM, WM)) < 0.2
    fr, mr = (2, 7, 1, 6), (30, 110, 17, 99)
    feats = [tok[s, 1:].reshape(gh, gw, D)[fr[0]:fr[1], fr[2]:fr[3]] for s in range(S)]
    masks = [mask[s, mr[0]:mr[1], mr[2]:mr[3]] for s in range(S)]
    res = (0.8, 0.8, 0.8)
    o = G.token_gather(feats, masks, res)
    t, src, cnt = ops.mask_gather(torch.from_numpy(tok.reshape(-1, D)).to(cuda), torch.from_numpy(mask.astype(np.uint8)).to(cuda),
                                  grid=(S, gh, gw, gh * gw + 1, 1), feat_roi=fr, mask_roi=mr, pe=dict(res=res))
    n = int(cnt.item())
    assert n == o["flat"].size and np.array_equal(src[:n].cpu().numpy(), o["src"])
    assert np.allclose(t[:n].cpu().numpy(), o["tokens"].astype(np.float32), rtol=0, atol=1e-6)


@pytest.mark.parametrize("D,gh,gw", [(48, 6, 7), (768, 5, 9), (96, 8, 8)])
def test_g1_pe_table_vector_path(cuda, D, gh, gw):
    """D % 24 == 0 takes the vectorised emit path (four 16-byte loads from the per-coordinate f64 PE table per 8 columns);
    non-square grids exercise the reference's xy-meshgrid quirk in the table row selection."""
    from oracle import gather_np as G
    from vit_deep_radiomics_b200 import ops
    rng = np.random.default_rng(100 + D)
    S, HM, WM = 7, 40, 56
    feats = rng.standard_normal((S, gh, gw, D)).astype(np.float32)
    masks = rng.random((S, HM, WM)) < 0.35
    res, noise = (0.8, 0.7, 1.25), (1.5, -2.0, 0.25)
    o = G.token_gather(list(feats), list(masks), res, noise)
    t, src, cnt = ops.mask_gather(torch.from_numpy(feats).to(cuda), torch.from_numpy(masks.astype(np.uint8)).to(cuda),
                                  pe=dict(res=res, noise=noise))
    n = int(cnt.item())
    assert n == o["flat"].size and np.array_equal(src[:n].cpu().numpy(), o["src"])
    want = o["tokens"].astype(np.float32)
    assert np.allclose(t[:n].cpu().numpy(), want, rtol=0, atol=2.5e-7 * max(1.0, np.abs(want).max()))


def test_g1_cap_and_bf16(cuda):
    from vit_deep_radiomics_b200 import ops
    rng = np.random.default_rng(9)
    feats = torch.from_numpy(rng.standard_normal((4, 10, 10, 32)).astype(np.float32)).to(cuda)
    mask = torch.from_numpy((rng.random((4, 40, 40)) < 0.5).astype(np.uint8)).to(cuda)
    full, src, cnt = ops.mask_gather(feats, mask)
    n = int(cnt.item())
    small, src2, cnt2 = ops.mask_gather(feats, mask, cap=7)     # count still reports the true total
    assert int(cnt2.item()) == n and torch.equal(small[:7], full[:7]) and torch.equal(src2[:7], src[:7])
    b16, _, _ = ops.mask_gather(feats.bfloat16(), mask)
    assert torch.equal(b16[:n], full[:n].bfloat16().float())


@pytest.mark.parametrize("name", ["sq", "rect", "edge", "empty"])
def test_g2_golden(cuda, golden_dir, name):
    from vit_deep_radiomics_b200 import create_pointcloud_dataframe as pcd
    g = np.load(os.path.join(golden_dir, "pointcloud_g2.npz"))
    img, res = g[f"{name}__img"], g[f"{name}__res"]
    mask = g[f"{name}__mask"].reshape(img.shape)          # stored flattened (the DataFrame column)
    df = pcd.to_pointcloud_df(img, mask, 1, res)
    for col in ("x", "y", "z", "raw", "mask", "mask_box"):
        assert np.array_equal(df[col].values, g[f"{name}__{col}"]), col
    box = pcd.pointcloud_box(img, mask, res, centre=False)
    keep = g[f"{name}__mask_box"]
    for col in ("x", "y", "z", "raw", "mask"):
        assert np.array_equal(box[col].values, g[f"{name}__{col}"][keep]), col


def test_g2_full_size_properties(cuda):
    """512x512x120 (BASELINE config C2): count = box volume, ascending flat order, payload == fancy index."""
    from vit_deep_radiomics_b200 import ops, synth
    img, mask, res, _ = synth.make_case("C2")
    i_t, m_t = torch.from_numpy(img).to(cuda), torch.from_numpy(mask.view(np.uint8)).to(cuda)
    bb = ops.voxel_bbox(m_t).cpu().numpy()
    rows, cols, sl = np.nonzero(mask.any(2).any(1))[0], np.nonzero(mask.any(2).any(0))[0], np.nonzero(mask.any(0).any(0))[0]
    assert bb.tolist() == [cols[0], cols[-1], rows[0], rows[-1], sl[0], sl[-1]]       # H == W: xi = col, yi = row
    cap = int((bb[1] - bb[0] + 1) * (bb[3] - bb[2] + 1) * (bb[5] - bb[4] + 1))
    flat, raw, mk, cnt = ops.voxel_gather(i_t, m_t, ops.voxel_bbox(m_t), cap)
    assert int(cnt.item()) == cap
    flat = flat.cpu().numpy().astype(np.int64)
    assert np.all(np.diff(flat) > 0)
    assert np.array_equal(raw.cpu().numpy(), img.reshape(-1)[flat]) and np.array_equal(mk.cpu().numpy().astype(bool), mask.reshape(-1)[flat])
    assert int(mk.sum().item()) == int(mask.sum())                                     # every in-mask voxel is inside the box


def test_g1_many_tiles_per_block(cuda):
    """A candidate grid large enough that one block of the cooperative kernel owns more tiles than it keeps ballots for
    (kKeep = 16): the recompute path of the rank phase must give the same stable order."""
    from oracle import gather_np as G
    from vit_deep_radiomics_b200 import ops
    rng = np.random.default_rng(21)
    S, gh, gw, D = 64, 200, 200, 8              # 2.56 M candidates = 1250 tiles of 2048 over <= 592 co-resident blocks... x17+ with a small grid
    feats = rng.standard_normal((S, gh, gw, D)).astype(np.float32)
    masks = rng.random((S, gh, gw)) < 0.03
    t, src, cnt = ops.mask_gather(torch.from_numpy(feats).to(cuda), torch.from_numpy(masks.astype(np.uint8)).to(cuda))
    n = int(cnt.item())
    sel = np.moveaxis(masks, 0, -1).reshape(-1)                # (h, w, S) flatten order
    assert n == int(sel.sum())
    flat = np.nonzero(sel)[0]
    s = src[:n].cpu().numpy().astype(np.int64)
    assert np.array_equal(s[:, 1] * (gw * S) + s[:, 2] * S + s[:, 0], flat)
    assert np.array_equal(t[:n].cpu().numpy(), np.moveaxis(feats, 0, 2).reshape(-1, D)[flat])


def test_g1_count_scan_and_table_slots(cuda):
    """The sharded-extraction contract: vdr_mask_count == the gather's count; offsets = device exclusive scan; three
    'patients' gathered into ONE table at device-resident row offsets == the concatenation of their own gathers, with the
    patient index in column 0 of the 4-column key."""
    from vit_deep_radiomics_b200 import ops
    rng = np.random.default_rng(33)
    D, pats = 48, []
    for p, (S, gh, gw, HM, WM, dens) in enumerate([(5, 6, 7, 30, 41, 0.4), (3, 9, 4, 27, 16, 0.0), (8, 5, 5, 20, 20, 0.7), (2, 4, 6, 16, 18, 0.2)]):
        feats = torch.from_numpy(rng.standard_normal((S, gh, gw, D)).astype(np.float32)).to(cuda)
        mask = torch.from_numpy((rng.random((S, HM, WM)) < dens).astype(np.uint8)).to(cuda)
        pats.append((feats, mask, dict(res=(0.8, 0.8, 1.0), noise=(0.1 * p, 0.0, -0.2))))
    counts = torch.zeros(len(pats), dtype=torch.int64, device=cuda)
    singles = []
    for p, (f, m, pe) in enumerate(pats):
        ops.mask_count(m, grid=f.shape[:3], out=counts[p:p + 1])
        t, s, c = ops.mask_gather(f, m, pe=pe)
        n = int(c.item())
        assert int(counts[p].item()) == n
        singles.append((t[:n].clone(), s[:n].clone()))
    offsets = ops.exclusive_scan_i64(counts)
    want_off = np.concatenate([[0], np.cumsum(counts.cpu().numpy())])
    assert np.array_equal(offsets.cpu().numpy(), want_off) and int(counts[1].item()) == 0      # an empty patient in the middle
    total = int(want_off[-1])
    table = dict(tokens=torch.full((total + 5, D), -7.0, device=cuda), src=torch.full((total + 5, 4), -7, dtype=torch.int32, device=cuda))
    for p in (2, 0, 3, 1):                                                                       # any order: slots are disjoint
        f, m, pe = pats[p]
        _, _, c = ops.mask_gather(f, m, pe=pe, table=dict(table, row_offset=offsets[p:p + 1], patient=p))
        assert int(c.item()) == int(counts[p].item())
    assert torch.equal(table["tokens"][:total], torch.cat([t for t, _ in singles]))
    keys = torch.cat([torch.cat([torch.full((s.shape[0], 1), p, dtype=torch.int32, device=cuda), s], 1) for p, (_, s) in enumerate(singles)])
    assert torch.equal(table["src"][:total], keys)
    assert (table["tokens"][total:] == -7.0).all() and (table["src"][total:] == -7).all()        # nothing beyond the slots
    # a large scan (more elements than threads)
    big = torch.from_numpy(rng.integers(0, 1 << 33, 5000)).to(cuda)
    assert np.array_equal(ops.exclusive_scan_i64(big).cpu().numpy(), np.concatenate([[0], np.cumsum(big.cpu().numpy())]))
