"""TEST INFRASTRUCTURE ONLY -- CPU oracle for ROC-AUC (and PRC / average precision).

numpy restatement of scikit-learn 1.9.0 `roc_curve` + `auc` exactly as the reference calls them
at src/eoe/training/ad_trainer.py:453-454 (train) and :517-518 (test):
    fpr, tpr, thresholds = roc_curve(labels, scores); auc = sklearn.metrics.auc(fpr, tpr)
and of `precision_recall_curve` / `average_precision_score` (ad_trainer.py:520-521).

scikit-learn is a third-party dependency of the reference (src/requirements.txt:9,
`scikit-learn>=1.0.2`, not pinned; 1.9.0 + numpy 2.3.5 in this image), so the algorithm is
restated from its published behaviour (sklearn/metrics/_ranking.py: _binary_clf_curve,
roc_curve(drop_intermediate=True), auc -> np.trapezoid) and the restatement is pinned
bit-for-bit against sklearn itself in tests/test_oracle_auc.py (sklearn ships in the image on
both boxes) and against tests/golden/auc_*.npz.

The only subtle part is the float64 *summation order*: np.trapezoid ends in ndarray.sum(),
which uses numpy's pairwise summation tree.  `pairwise_sum` spells that tree out; it is the
shape the CUDA reduction must reproduce.
"""
import numpy as np


def pairwise_sum(a):
    """numpy's pairwise summation (numpy/_core/src/umath/loops_utils.h.src, *_pairwise_sum)
    for a contiguous float64 vector: n<8 serial; n<=128 eight strided accumulators combined
    ((r0+r1)+(r2+r3))+((r4+r5)+(r6+r7)) then the n%8 tail serially; else split at n/2 rounded
    down to a multiple of 8 and recurse."""
    a = np.asarray(a, dtype=np.float64)
    n = a.shape[0]
    if n < 8:
        # numpy starts from -0.0 so that the sum of nothing / of -0.0s keeps its sign
        res = np.float64(-0.0)
        for i in range(n):
            res = res + a[i]
        return np.float64(res)
    if n <= 128:
        r = [np.float64(a[j]) for j in range(8)]
        i = 8
        while i < n - (n % 8):
            for j in range(8):
                r[j] = r[j] + a[i + j]
            i += 8
        res = ((r[0] + r[1]) + (r[2] + r[3])) + ((r[4] + r[5]) + (r[6] + r[7]))
        while i < n:
            res = res + a[i]
            i += 1
        return np.float64(res)
    h = n // 2
    h -= h % 8
    return np.float64(pairwise_sum(a[:h]) + pairwise_sum(a[h:]))


def binary_clf_curve(labels, scores):
    """sklearn _binary_clf_curve: stable descending sort, distinct-value indices, tps/fps."""
    y = np.asarray(labels)
    s = np.asarray(scores)
    if not np.all(np.isfinite(s.astype(np.float64))):
        raise ValueError("Input contains NaN or infinity.")
    y = (y == 1)
    order = np.argsort(s, kind="mergesort")[::-1]
    s = s[order]
    y = y[order]
    distinct = np.where(np.diff(s))[0]
    idx = np.r_[distinct, y.size - 1]
    tps = np.cumsum(y, dtype=np.float64)[idx]
    fps = 1 + idx - tps
    return fps, tps, s[idx]


def roc_curve(labels, scores):
    """sklearn roc_curve(y, s) with default drop_intermediate=True. Returns fpr, tpr, thresholds."""
    fps, tps, thr = binary_clf_curve(labels, scores)
    if len(fps) > 2:
        keep = np.where(np.r_[True, np.logical_or(np.diff(fps, 2), np.diff(tps, 2)), True])[0]
        fps, tps, thr = fps[keep], tps[keep], thr[keep]
    tps = np.r_[0, tps]
    fps = np.r_[0, fps]
    thr = np.r_[np.inf, thr]
    fpr = fps / fps[-1] if fps[-1] > 0 else np.repeat(np.nan, fps.shape)
    tpr = tps / tps[-1] if tps[-1] > 0 else np.repeat(np.nan, tps.shape)
    return fpr, tpr, thr


def trapezoid_terms(fpr, tpr):
    """np.trapezoid(tpr, fpr) integrand: diff(x) * (y[1:] + y[:-1]) / 2.0, evaluated in that order."""
    return np.diff(fpr) * (tpr[1:] + tpr[:-1]) / 2.0


def auc(fpr, tpr):
    """sklearn.metrics.auc(fpr, tpr) for monotone-increasing fpr (direction = +1)."""
    return float(pairwise_sum(trapezoid_terms(fpr, tpr)))


def roc_auc(labels, scores):
    """The number the reference keeps: auc(*roc_curve(labels, scores)[:2])."""
    fpr, tpr, _ = roc_curve(labels, scores)
    return auc(fpr, tpr)


def precision_recall_curve(labels, scores):
    """sklearn precision_recall_curve (default drop_intermediate=False)."""
    fps, tps, thr = binary_clf_curve(labels, scores)
    ps = tps + fps
    precision = np.zeros_like(tps)
    np.divide(tps, ps, out=precision, where=(ps != 0))
    recall = np.ones_like(tps) if tps[-1] == 0 else tps / tps[-1]
    sl = slice(None, None, -1)
    return np.hstack((precision[sl], 1)), np.hstack((recall[sl], 0)), thr[sl]


def average_precision(labels, scores):
    """sklearn average_precision_score (binary): max(0, -sum(diff(recall) * precision[:-1]))."""
    p, r, _ = precision_recall_curve(labels, scores)
    return float(max(0.0, -pairwise_sum(np.diff(r) * np.array(p)[:-1])))
