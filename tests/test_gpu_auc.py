"""GPU ROC-AUC / PRC (through the C ABI) must be BIT-EXACT with the oracle restatement of sklearn
(oracle/auc.py), the golden fixtures generated with sklearn, and sklearn itself (ships in the image)."""
import os

import numpy as np
import pytest
import torch

from oracle import auc as oauc
from oracle import golden_inputs as gi

pytestmark = pytest.mark.gpu
DEV = "cuda"


def _run(y, s, with_prc=True):
    from eoe_b200 import metrics
    st = torch.from_numpy(np.ascontiguousarray(s)).to(DEV)
    yt = torch.from_numpy(np.ascontiguousarray(y)).to(DEV)
    return metrics.roc_curve_auc(st, yt, with_prc=with_prc)


@pytest.mark.parametrize("name", gi.AUC_CASES)
def test_auc_golden_bit_exact(golden_dir, name):
    g = np.load(os.path.join(golden_dir, f"auc_{name}.npz"))
    y, s = gi.auc_inputs(name)
    roc, prc = _run(y, s)
    assert roc.auc == float(g["auc"])
    assert np.array_equal(roc.fpr, g["fpr"]) and np.array_equal(roc.tpr, g["tpr"])
    assert np.array_equal(roc.ths.astype(np.float64), g["thresholds"])
    assert prc.avg_prec == float(g["ap"])
    assert np.array_equal(prc.prec, g["precision"]) and np.array_equal(prc.rec, g["recall"])


@pytest.mark.parametrize("n", [2, 3, 7, 8, 9, 127, 128, 129, 130, 1000, 4095, 4096, 4097, 10000, 65536, 100003])
@pytest.mark.parametrize("kind", ["f32", "f16", "coarse", "signed"])
def test_auc_vs_oracle_bit_exact(n, kind):
    rng = np.random.default_rng(n * 13 + len(kind))
    s = (1 - np.exp(-np.abs(rng.standard_normal(n)))).astype(np.float32)
    if kind == "f16":
        s = s.astype(np.float16)
    elif kind == "coarse":
        s = (np.round(s * 20) / np.float32(20)).astype(np.float32)
    elif kind == "signed":
        s = rng.standard_normal(n).astype(np.float32) * np.float32(1e3)
    y = (rng.random(n) < 0.35).astype(np.int64)
    y[0], y[-1] = 1, 0
    roc, prc = _run(y, s)
    fpr, tpr, thr = oauc.roc_curve(y, s)
    assert np.array_equal(roc.fpr, fpr) and np.array_equal(roc.tpr, tpr)
    assert np.array_equal(roc.ths.astype(np.float64), thr.astype(np.float64))
    assert roc.auc == oauc.auc(fpr, tpr)
    assert prc.avg_prec == oauc.average_precision(y, s)


@pytest.mark.parametrize("n,kind", [(1_000_000, "f32"), (1_000_000, "f16"), (1_000_000, "distinct1000"), (3_000_017, "f32")])
def test_auc_1m_bit_exact_vs_sklearn(n, kind):
    """BASELINE size (1 M scores) against sklearn itself, the reference's AUC (ad_trainer.py:8)."""
    from sklearn.metrics import auc, average_precision_score, roc_curve
    rng = np.random.default_rng(0)
    s = (1 - np.exp(-np.abs(rng.standard_normal(n)))).astype(np.float32)
    if kind == "f16":
        s = s.astype(np.float16).astype(np.float32)
    elif kind == "distinct1000":
        s = (np.floor(s * 1000) / np.float32(1000)).astype(np.float32)
    y = (rng.random(n) < 0.5).astype(np.int64)
    fpr, tpr, thr = roc_curve(y, s)
    roc, prc = _run(y, s)
    assert roc.auc == auc(fpr, tpr)
    assert np.array_equal(roc.fpr, fpr) and np.array_equal(roc.tpr, tpr)
    assert np.array_equal(roc.ths[1:], thr[1:])
    assert prc.avg_prec == average_precision_score(y, s)


def test_auc_permutation_invariance_and_reuse():
    from eoe_b200 import metrics
    rng = np.random.default_rng(3)
    n = 200_000
    s = torch.from_numpy(rng.standard_normal(n).astype(np.float32)).to(DEV)
    y = torch.from_numpy((rng.random(n) < 0.2).astype(np.int64)).to(DEV)
    a = metrics.roc_auc(s, y)
    p = torch.randperm(n, device=DEV)
    assert metrics.roc_auc(s[p], y[p]) == a          # sort + tie merge make the order irrelevant
    assert metrics.roc_auc(s, y) == a                # workspace reuse leaves no state behind
    assert abs(metrics.roc_auc(-s, y) - (1 - a)) < 1e-12


def test_auc_ignore_negative_labels_and_errors():
    from eoe_b200 import metrics
    rng = np.random.default_rng(4)
    n = 5000
    s = rng.standard_normal(n).astype(np.float32)
    y = rng.integers(-1, 2, n).astype(np.int64)          # -1 = unlabeled (datasets/custom.py:17)
    keep = y >= 0
    st, yt = torch.from_numpy(s).to(DEV), torch.from_numpy(y).to(DEV)
    got = metrics.roc_auc(st, yt, ignore_negative_labels=True)
    assert got == oauc.roc_auc(y[keep], s[keep])
    s2 = s.copy(); s2[17] = np.nan
    with pytest.raises(ValueError):
        metrics.roc_auc(torch.from_numpy(s2).to(DEV), torch.from_numpy((y > 0).astype(np.int64)).to(DEV))
    roc, prc = metrics.roc_curve_auc(st, torch.ones(n, dtype=torch.int64, device=DEV))
    assert roc is None and prc is None                   # ad_trainer.py:516,523-527: single class -> None
