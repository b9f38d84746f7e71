"""CPU: the oracle restatements reproduce the fixtures generated from the LIVE reference
(oracle/make_golden.py).  Tolerances: heads 1e-5 relative (fp32 summation order only),
AUC / curves bit-exact, ViT features 2e-5 absolute on O(1) values."""
import os

import numpy as np
import pytest
import torch

from oracle import auc as oauc
from oracle import golden_inputs as gi
from oracle import heads as oh
from oracle import vit as ovit


def _load(golden_dir, name):
    return np.load(os.path.join(golden_dir, name))


def _close(a, b, rtol=1e-5, atol=1e-7):
    np.testing.assert_allclose(np.asarray(a, np.float64), np.asarray(b, np.float64), rtol=rtol, atol=atol)


def test_hsc_oracle_matches_reference_fixture(golden_dir):
    g = _load(golden_dir, "heads.npz")
    z, y = gi.hsc_inputs()
    assert float(g["hsc_in_sum"]) == float(z.astype(np.float64).sum())
    _close(oh.hsc_score(z), g["hsc_score"])
    for nom in (0, 1):
        _close(oh.hsc_loss(z, y, nom), g[f"hsc_loss_nom{nom}"])
        _close(oh.hsc_grad(z, y, nom), g[f"hsc_grad_nom{nom}"], rtol=2e-5, atol=1e-9)


def test_bce_oracle_matches_reference_fixture(golden_dir):
    g = _load(golden_dir, "heads.npz")
    x, y = gi.bce_inputs()
    assert float(g["bce_in_sum"]) == float(x.astype(np.float64).sum())
    _close(oh.bce_loss(x, y), g["bce_loss"])
    _close(oh.bce_grad(x, y).reshape(-1, 1), g["bce_grad"], atol=1e-10)
    for nom in (0, 1):
        _close(oh.bce_score(x, nom), g[f"bce_score_nom{nom}"], atol=1e-12)


def test_dsad_dsvdd_focal_oracle_matches_reference_fixture(golden_dir):
    """dsad.py:13-22, dsvdd.py:23-27, focal.py:11-39 run from the live reference (oracle/make_golden.py)."""
    g = _load(golden_dir, "heads.npz")
    z, y = gi.hsc_inputs()
    _close(oh.dsad_score(z), g["dsad_score"])
    for nom in (0, 1):
        _close(oh.dsad_loss(z, y, nom), g[f"dsad_loss_nom{nom}"])
        _close(oh.dsad_grad(z, y, nom), g[f"dsad_grad_nom{nom}"], rtol=2e-5, atol=1e-9)
    zd, cd = gi.dsvdd_inputs()
    assert float(g["dsvdd_in_sum"]) == float(zd.astype(np.float64).sum() + cd.astype(np.float64).sum())
    _close(oh.dsvdd_score(zd, cd), g["dsvdd_score"])
    _close(oh.dsvdd_loss(zd, cd), g["dsvdd_loss"])
    _close(oh.dsvdd_grad(zd, cd), g["dsvdd_grad"], atol=1e-10)
    x, yb = gi.bce_inputs()
    _close(oh.focal_loss(x, yb), g["focal_loss"])
    _close(oh.focal_grad(x, yb).reshape(-1, 1), g["focal_grad"], rtol=2e-5, atol=1e-10)
    for nom in (0, 1):
        _close(oh.focal_score(x, nom), g[f"focal_score_nom{nom}"], atol=1e-12)


@pytest.mark.parametrize("K", [2, 10, 30])
def test_clip_oracle_matches_reference_fixture(golden_dir, K):
    g = _load(golden_dir, "heads.npz")
    z, y, c = gi.clip_inputs(K)
    assert float(g[f"clip_in_sum_K{K}"]) == float(z.astype(np.float64).sum() + c.astype(np.float64).sum())
    _close(oh.clip_score(z, c), g[f"clip_score_K{K}"], rtol=2e-4, atol=1e-30)
    for mode in ("one_vs_rest", "leave_one_out"):
        for nom in (0, 1):
            loo = mode == "leave_one_out"
            _close(oh.clip_oe_loss(z, y, c, nom, loo), g[f"clip_loss_K{K}_{mode}_nom{nom}"], rtol=2e-5)
            _close(oh.clip_oe_grad(z, y, c, nom, loo), g[f"clip_grad_K{K}_{mode}_nom{nom}"], rtol=2e-3, atol=2e-7)


@pytest.mark.parametrize("name", gi.AUC_CASES)
def test_auc_oracle_bit_exact_vs_fixture(golden_dir, name):
    g = _load(golden_dir, f"auc_{name}.npz")
    y, s = gi.auc_inputs(name)
    assert float(g["in_sum"]) == float(s.astype(np.float64).sum() + y.sum())
    fpr, tpr, thr = oauc.roc_curve(y, s)
    assert np.array_equal(fpr, g["fpr"]) and np.array_equal(tpr, g["tpr"])
    assert np.array_equal(thr.astype(np.float64), g["thresholds"])
    assert oauc.auc(fpr, tpr) == float(g["auc"])
    assert oauc.average_precision(y, s) == float(g["ap"])
    p, r, _ = oauc.precision_recall_curve(y, s)
    assert np.array_equal(p, g["precision"]) and np.array_equal(r, g["recall"])


@pytest.mark.parametrize("n", [2, 7, 8, 9, 127, 128, 129, 1000, 4099, 100003])
@pytest.mark.parametrize("kind", ["f32", "f16", "coarse"])
def test_auc_oracle_bit_exact_vs_sklearn(n, kind):
    """sklearn (the reference's third-party AUC, ad_trainer.py:8) ships in the image on both boxes."""
    from sklearn.metrics import auc, roc_curve
    rng = np.random.default_rng(n * 7 + len(kind))
    s = (1 - np.exp(-np.abs(rng.standard_normal(n)))).astype(np.float32)
    if kind == "f16":
        s = s.astype(np.float16)
    elif kind == "coarse":
        s = np.round(s * 20) / np.float32(20)
    y = (rng.random(n) < 0.35).astype(np.int64)
    y[0], y[-1] = 1, 0
    fpr, tpr, thr = roc_curve(y, s)
    f2, t2, th2 = oauc.roc_curve(y, s)
    assert np.array_equal(fpr, f2) and np.array_equal(tpr, t2)
    assert np.array_equal(thr.astype(np.float64), th2.astype(np.float64))
    assert auc(fpr, tpr) == oauc.auc(f2, t2)


def test_pairwise_sum_is_numpy_sum():
    rng = np.random.default_rng(5)
    for n in [0, 1, 5, 7, 8, 9, 15, 16, 127, 128, 129, 255, 256, 1000, 4097, 99999]:
        a = rng.standard_normal(n) * 10.0 ** rng.integers(-8, 8, n)
        assert oauc.pairwise_sum(a) == a.sum(), n


def test_auc_nonfinite_raises():
    with pytest.raises(ValueError):
        oauc.roc_curve(np.array([0, 1, 1]), np.array([0.1, np.nan, 0.3], np.float32))


@pytest.mark.parametrize("patch", [32, 16])
def test_vit_oracle_matches_reference_fixture(golden_dir, patch):
    g = _load(golden_dir, f"vit_b{patch}.npz")
    sd = ovit.synth_state_dict(patch, seed=gi.VIT_WEIGHT_SEED)
    imgs = gi.vit_images()
    assert abs(float(g["img_sum"]) - float(imgs.double().sum())) < 1e-9
    assert abs(float(g["w_sum"]) - float(sum(v.double().sum() for v in sd.values()))) < 1e-6
    torch.set_num_threads(max(1, os.cpu_count() or 1))
    feats = ovit.encode_image(sd, imgs).numpy()
    np.testing.assert_allclose(feats, g["features"], rtol=0, atol=2e-5)


def score_errors(scores, want):
    """relative score errors (north_star's measure): median, max, fraction within 1e-3."""
    rel = np.abs(scores.astype(np.float64) - want.astype(np.float64)) / np.abs(want.astype(np.float64))
    return float(np.median(rel)), float(rel.max()), float((rel <= 1e-3).mean())


@pytest.mark.parametrize("patch,K", gi.SCORE_PARITY_CFGS)
def test_end_to_end_score_fixture(golden_dir, patch, K):
    """score_parity_b*.npz (live reference: encode_image -> ADClipTrainer.compute_anomaly_score on 64 images, cfg2 / cfg3):
    (a) the head oracle reproduces the fixture's scores from the fixture's features (1e-3 relative, far-tail scores);
    (b) the fp32 encoder oracle reproduces them END TO END on a 16-image slice (kept small: CPU time);
    (c) the reference's own GPU precision (half run of the live reference) is itself 3e-3 ... 5e-3 (median) away from its
        fp32 answer on these scores -- the yardstick tests/test_gpu_encoder.py::test_end_to_end_scores holds us to."""
    g = _load(golden_dir, f"score_parity_b{patch}.npz")
    imgs, text, labels = gi.score_parity_inputs(K)
    assert abs(float(g["img_sum"]) - float(imgs.double().sum())) < 1e-9
    assert np.array_equal(g["text"], text) and np.array_equal(g["labels"], labels)
    np.testing.assert_allclose(oh.clip_score(g["features"], text), g["scores"], rtol=1e-3, atol=1e-30)
    sd = ovit.synth_state_dict(patch, seed=gi.VIT_WEIGHT_SEED)
    torch.set_num_threads(max(1, os.cpu_count() or 1))
    feats = ovit.encode_image(sd, imgs[:16]).numpy()
    med, mx, _ = score_errors(oh.clip_score(feats, text), g["scores"][:16])
    assert mx < 1e-3, (med, mx)                       # fp32 oracle end to end: within north_star's 1e-3 on every score
    med_h, mx_h, frac_h = score_errors(g["scores_ref_half"], g["scores"])
    assert 2e-3 < med_h < 8e-3 and frac_h < 0.3, (med_h, mx_h, frac_h)
    assert oauc.roc_auc(labels, g["scores"]) == oauc.roc_auc(labels, g["scores_ref_half"])


def test_split_operand_emulation_meets_the_score_bar(golden_dir):
    """The precise mode's emulation (oracle.vit F16X2: fp16 (hi, lo) pairs at every stored 16-bit tensor, the softmax
    probabilities included) against the live reference's fp32 scores, cfg2, 16-image slice: every score within 1e-4
    (measured max 2.6e-5, features 2.4e-6 -- the fp32 reference's own rounding noise), where the single-fp16 emulation has
    half of them outside 1e-3 -- the CPU statement of what tests/test_gpu_encoder_split.py::test_end_to_end_scores_split
    asserts for the kernels on all 64 images."""
    patch, K = gi.SCORE_PARITY_CFGS[0]
    g = _load(golden_dir, f"score_parity_b{patch}.npz")
    imgs, text, _ = gi.score_parity_inputs(K)
    sd = ovit.synth_state_dict(patch, seed=gi.VIT_WEIGHT_SEED)
    torch.set_num_threads(max(1, os.cpu_count() or 1))
    f2 = ovit.encode_image(sd, imgs[:16], operand_dtype=ovit.F16X2, fold_layernorm=True).numpy()
    f1 = ovit.encode_image(sd, imgs[:16], operand_dtype=torch.float16, fold_layernorm=True).numpy()
    med2, mx2, frac2 = score_errors(oh.clip_score(f2, text), g["scores"][:16])
    med1, mx1, frac1 = score_errors(oh.clip_score(f1, text), g["scores"][:16])
    assert mx2 < 1e-4 and frac2 == 1.0, (med2, mx2)
    assert med2 < 0.1 * med1, (med2, med1)
    want = g["features"][:16]
    assert np.linalg.norm(f2 - want) / np.linalg.norm(want) < 1e-5


def test_text_oracle_matches_reference_fixture(golden_dir):
    from oracle import text as otext
    g = _load(golden_dir, "text.npz")
    sd = otext.synth_text_state_dict(seed=gi.TEXT_WEIGHT_SEED)
    tokens = gi.text_tokens()
    assert int(g["tok_sum"]) == int(tokens.sum())
    assert abs(float(g["w_sum"]) - float(sum(v.double().sum() for v in sd.values()))) < 1e-6
    torch.set_num_threads(max(1, os.cpu_count() or 1))
    feats = otext.encode_text(sd, tokens).numpy()
    np.testing.assert_allclose(feats, g["features"], rtol=0, atol=2e-5)


def test_vit_flop_counts():
    assert abs(ovit.flops_per_image(32) / 1e9 - 8.818) < 1e-3
    assert abs(ovit.flops_per_image(16) / 1e9 - 35.127) < 1e-3
