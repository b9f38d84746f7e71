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


def test_auc_sort_tile_switch_and_grown_workspace():
    """The sort uses 2048-key tiles up to 655 360 scores and 4096-key tiles above: both sides of the switch against sklearn,
    the smaller size through a workspace that was sized for the larger one (grow-only reuse)."""
    from sklearn.metrics import roc_auc_score
    from eoe_b200 import metrics
    rng = np.random.default_rng(11)
    ws = metrics.AucWorkspace()
    for n in (700_001, 655_361, 655_360, 300_000):
        s = (1 - np.exp(-np.abs(rng.standard_normal(n)))).astype(np.float32)
        y = (rng.random(n) < 0.3).astype(np.int64)
        out, info, _ = metrics.roc_auc_device(torch.from_numpy(s).to(DEV), torch.from_numpy(y).to(DEV), workspace=ws)
        assert int(info[4].item()) == 0
        assert out[0].item() == roc_auc_score(y, s), n
    assert ws.n == 700_001


@pytest.mark.parametrize("n", [10_000, 200_000, 1_000_000])
def test_auc_replays_from_a_cuda_graph(n):
    """The whole call -- one launch (n = 10 000), or the memset + ten kernels of the tiled pipeline chained by programmatic
    dependent launch -- is captured once and replayed on new scores: bit-identical to sklearn every time (the dependent-launch
    edges must survive capture, and the workspace must carry no state from one replay to the next)."""
    from sklearn.metrics import roc_auc_score
    from eoe_b200 import metrics
    rng = np.random.default_rng(n)
    s_static = torch.empty(n, dtype=torch.float32, device=DEV)
    y_static = torch.empty(n, dtype=torch.int64, device=DEV)
    ws = metrics.AucWorkspace()

    def fresh():
        s = (1 - np.exp(-np.abs(rng.standard_normal(n)))).astype(np.float32)
        y = (rng.random(n) < 0.35).astype(np.int64)
        s_static.copy_(torch.from_numpy(s))
        y_static.copy_(torch.from_numpy(y))
        return s, y

    fresh()
    side = torch.cuda.Stream()
    side.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(side):
        metrics.roc_auc_device(s_static, y_static, workspace=ws)
    torch.cuda.current_stream().wait_stream(side)
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        out, info, _ = metrics.roc_auc_device(s_static, y_static, workspace=ws)
    for _ in range(3):
        s, y = fresh()
        g.replay()
        torch.cuda.synchronize()
        assert int(info[4].item()) == 0
        assert out[0].item() == roc_auc_score(y, s)


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


@pytest.mark.parametrize("n", [2, 5, 9, 100, 1023, 1024, 1025, 3000, 8191, 10000, 12288, 12289, 16383, 16384, 16385, 24577,
                               49152, 49153, 65536, 100003, 131071, 131072])
@pytest.mark.parametrize("kind", ["f32", "f16", "coarse", "signed", "unlabeled"])
def test_auc_single_launch_equals_tiled_pipeline(n, kind):
    """n <= EOE_AUC_SINGLE_LAUNCH_MAX runs as ONE kernel launch: one CTA up to 12 288 scores (the sizes the reference
    evaluates, ad_trainer.py:452-455,516-522), a cluster of 8 CTAs sorting through distributed shared memory above.  All
    three paths (one CTA, cluster -- forced up to its 131 072 limit --, tiled multi-kernel pipeline) must give the same
    bits for every output."""
    from eoe_b200 import _lib, metrics
    rng = np.random.default_rng(n * 17 + len(kind))
    s = (1 - np.exp(-np.abs(rng.standard_normal(n)))).astype(np.float32)
    y = (rng.random(n) < 0.35).astype(np.int64)
    y[0], y[-1] = 1, 0
    if kind == "f16":
        s = s.astype(np.float16)
    elif kind == "coarse":
        s = (np.round(s * 20) / np.float32(20)).astype(np.float32)
    elif kind == "signed":
        s = rng.standard_normal(n).astype(np.float32) * np.float32(1e3)
    elif kind == "unlabeled" and n > 5:
        y[2::5] = -1
    st, yt = torch.from_numpy(s).to(DEV), torch.from_numpy(y).to(DEV)
    ign = kind == "unlabeled"
    launches0 = _lib.lib().eoe_launch_count()
    a = metrics.roc_curve_auc(st, yt, with_prc=True, ignore_negative_labels=ign)            # the default path for this n
    if n <= _lib.EOE_AUC_SINGLE_LAUNCH_MAX:
        assert _lib.lib().eoe_launch_count() - launches0 == 1
    launches0 = _lib.lib().eoe_launch_count()
    arms = [metrics.roc_curve_auc(st, yt, with_prc=True, ignore_negative_labels=ign, force_tiled=True)]
    assert _lib.lib().eoe_launch_count() - launches0 > 10
    launches0 = _lib.lib().eoe_launch_count()
    arms.append(metrics.roc_curve_auc(st, yt, with_prc=True, ignore_negative_labels=ign, force_cluster=True))
    assert _lib.lib().eoe_launch_count() - launches0 == 1
    if n <= 16384:
        launches0 = _lib.lib().eoe_launch_count()
        arms.append(metrics.roc_curve_auc(st, yt, with_prc=True, ignore_negative_labels=ign, force_single_cta=True))
        assert _lib.lib().eoe_launch_count() - launches0 == 1
    for b in arms:
        for x, w in zip(a, b):
            for f in ("auc", "avg_prec"):
                if hasattr(x, f):
                    assert getattr(x, f) == getattr(w, f)
            for f in ("tpr", "fpr", "prec", "rec", "ths"):
                if hasattr(x, f):
                    assert np.array_equal(getattr(x, f), getattr(w, f)), f
    keep = y >= 0
    assert a[0].auc == oauc.roc_auc(y[keep], s[keep])


@pytest.mark.parametrize("n", [50, 3000, 40000])
def test_prc_thresholds_like_sklearn_and_reference_logger(n):
    """PRC.ths = precision_recall_curve's thresholds (distinct scores, increasing): the reference's logger averages
    np.asarray(res.ths) over seeds and classes (utils/logger.py:103-111), so None would break its own aggregation."""
    from sklearn.metrics import precision_recall_curve
    rng = np.random.default_rng(n)
    s = (np.round((1 - np.exp(-np.abs(rng.standard_normal(n)))) * 500) / 500).astype(np.float32)
    y = (rng.random(n) < 0.4).astype(np.int64)
    roc, prc = _run(y, s)
    p, r, t = precision_recall_curve(y, s)
    assert np.array_equal(prc.ths, t) and prc.ths.dtype == t.dtype
    assert np.array_equal(prc.prec, p) and np.array_equal(prc.rec, r)
    # what mean_plot does with a list of curves (logger.py:103-118): index y, x and ths with picks from range(len(ths))
    for res in (roc, prc):
        ths = np.asarray(res.ths)
        pick = sorted(np.random.default_rng(0).choice(len(ths), size=min(len(ths), 20), replace=False))
        assert np.asarray(res.get_y())[pick].shape == np.asarray(res.get_x())[pick].shape == ths[pick].shape
