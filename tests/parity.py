"""Stated floating-point bounds of the parity tests (BASELINE.json north_star: "max-abs/rel error per descriptor and classifier
logit, plus CLS-token cosine >= 0.999 in BF16 versus the reference FP32").

Every check records its measured triplet (max |err|, rms relative error, min row cosine); a `-m gpu` session writes them to
``gpurun_out/parity_measured.json``.  BOUNDS holds, per key, about twice the value measured on a B200 (round 2;
``tests/golden/parity_measured_r02.json`` keeps the numbers of the last full `-m gpu` run of the round, on the final kernels; where two
runs of the round differed -- the attention kernels changed between them -- a bound is twice the larger value); keys without an entry
fall back to DEFAULT."""
import numpy as np

DEFAULT = dict(abs=0.12, rel=0.02, cos=0.999)
BOUNDS: dict = {
    # key: about 2 x (max |err|, rms relative error), 1 - 2 x (1 - min cosine) of the round-2 B200 run (tests/golden/parity_measured_r02.json)
    "bimodal gradients (golden, both)": dict(abs=0.00023, rel=0.024, cos=0.99987),
    "bimodal gradients (golden, ct)": dict(abs=0.0024, rel=0.047, cos=0.99948),
    "bimodal gradients (golden, pet)": dict(abs=0.0038, rel=0.039, cos=0.99966),
    "classifier CLS (d256 ff1024 h4 L2, n=2000)": dict(abs=0.045, rel=0.011, cos=0.99997),
    "classifier CLS (golden, small)": dict(abs=0.025, rel=0.009, cos=0.99998),
    "classifier gradients (golden, small)": dict(abs=0.0015, rel=0.02, cos=0.99992),
    "classifier logits (d256 ff1024 h4 L2, n=2000)": dict(abs=0.012, rel=0.013, cos=0.99998),
    "classifier logits (golden, small)": dict(abs=0.0021, rel=0.0061, cos=0.99999),
    "descriptors medsam (SAM ViT-B)@1024x1024": dict(abs=0.12, rel=0.023, cos=0.99982),
    "descriptors sam_tiny vs HF golden": dict(abs=0.063, rel=0.014, cos=0.99991),
    "descriptors sam_tiny@128x1024 (kernel variants)": dict(abs=0.062, rel=0.013, cos=0.99991),
    "descriptors sam_tiny@192x320": dict(abs=0.06, rel=0.014, cos=0.99990),
    "descriptors sam_tiny@224x224": dict(abs=0.06, rel=0.013, cos=0.99992),
    "descriptors sam_tiny@256x256": dict(abs=0.061, rel=0.013, cos=0.99992),
    "descriptors unfolded vit_s16@256": dict(abs=0.14, rel=0.019, cos=0.99988),
    "descriptors unfolded vit_t16@64": dict(abs=0.052, rel=0.009, cos=0.99996),
    "descriptors vit_b16@512x512 (C2)": dict(abs=0.15, rel=0.019, cos=0.99989),          # the bench line's workload
    "descriptors vit_l14@224x224 (C4)": dict(abs=0.24, rel=0.026, cos=0.99980),
    "descriptors vit_l14@56x56 B2": dict(abs=0.17, rel=0.025, cos=0.99982),
    "descriptors vit_s16@224x224 B8": dict(abs=0.14, rel=0.018, cos=0.99989),            # C1
    "descriptors vit_t16@48x80 B2": dict(abs=0.059, rel=0.0085, cos=0.99997),
    "descriptors vit_t16@64x64 B3": dict(abs=0.051, rel=0.0086, cos=0.99997),
    "folded vs unfolded LayerNorm vit_s16": dict(abs=0.16, rel=0.02, cos=0.99987),
    "folded vs unfolded LayerNorm vit_t16": dict(abs=0.055, rel=0.0056, cos=0.99996),
    "generate_features C1": dict(abs=0.15, rel=0.018, cos=0.99989),
    "generate_features T0": dict(abs=0.055, rel=0.0094, cos=0.99997),
    "generate_features resized crop": dict(abs=0.04, rel=0.0095, cos=0.99997),
    "get_dense_descriptor vit_t16@64x64": dict(abs=0.052, rel=0.0091, cos=0.99997),
    "point-cloud tokens C1": dict(abs=0.13, rel=0.018, cos=0.99990),
    "point-cloud tokens T0": dict(abs=0.052, rel=0.009, cos=0.99997),
    "point-cloud tokens vit_b16@512x512 (C2)": dict(abs=0.13, rel=0.019, cos=0.99990),
    "point-cloud tokens vit_l14@224x224 (C4)": dict(abs=0.24, rel=0.026, cos=0.99980),
}
LOG: dict = {}


def triplet(got, want):
    got, want = np.asarray(got, dtype=np.float64), np.asarray(want, dtype=np.float64)
    err = float(np.abs(got - want).max()) if got.size else 0.0
    den = float((want ** 2).sum())
    rel = float(np.sqrt(((got - want) ** 2).sum() / den)) if den > 0 else 0.0
    g2, w2 = got.reshape(-1, got.shape[-1]), want.reshape(-1, want.shape[-1])
    n = np.linalg.norm(g2, axis=1) * np.linalg.norm(w2, axis=1)
    cos = float(((g2 * w2).sum(1)[n > 0] / n[n > 0]).min()) if (n > 0).any() else 1.0
    return dict(abs=err, rel=rel, cos=cos)


def check(key, got, want, default=None):
    """Assert the triplet of (got, want) against BOUNDS[key]; rows = last axis (one descriptor / token / logit vector per row)."""
    t = triplet(got, want)
    prev = LOG.get(key)
    LOG[key] = t if prev is None else dict(abs=max(t["abs"], prev["abs"]), rel=max(t["rel"], prev["rel"]), cos=min(t["cos"], prev["cos"]))
    b = BOUNDS.get(key) or default or DEFAULT
    assert t["abs"] <= b["abs"] and t["rel"] <= b["rel"] and t["cos"] >= b["cos"], (key, t, b)
    return t
