"""TEST INFRASTRUCTURE ONLY -- CPU oracle for the loss / anomaly-score heads.

numpy restatement (float32 arithmetic, reference evaluation order) of the reference's
objective hooks.  Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline leg may
import this; the product path (eoe_b200/) never does.

Pinned against the live reference by tests/test_oracle_vs_reference.py (runs where
/root/reference is mounted) and against the committed fixtures in tests/golden/
(generated from the live reference by oracle/make_golden.py).

Reference lines restated (paths relative to /root/reference/src/eoe):
  hsc_*        training/hsc.py:12-21
  bce_*        training/bce.py:15-20   (torch binary_cross_entropy_with_logits, mean reduction)
  clip_score   training/clip.py:66-79
  clip_oe_loss training/clip.py:81-103
  dsad_*       training/dsad.py:13-22
  dsvdd_*      training/dsvdd.py:23-27
  focal_*      training/focal.py:11-39   (FocalLoss, gamma = 2, eps = 1e-7)
"""
import numpy as np

F32 = np.float32


def _f32(a):
    return np.ascontiguousarray(np.asarray(a, dtype=np.float32))


# --------------------------------------------------------------------------- HSC
def hsc_dists(z):
    """training/hsc.py:13,18  dists = sqrt(norm(z,2,dim=1)**2 + 1) - 1  (fp32)."""
    z = _f32(z)
    nrm = np.sqrt(np.sum(z * z, axis=1, dtype=np.float32)).astype(F32)
    return (np.sqrt(nrm * nrm + F32(1)) - F32(1)).astype(F32)


def hsc_score(z):
    """training/hsc.py:12-15  scores = 1 - exp(-dists)."""
    d = hsc_dists(z)
    return (F32(1) - np.exp(-d)).astype(F32)


def hsc_loss(z, labels, nominal_label=0):
    """training/hsc.py:17-21  mean(where(labels==nominal, dists, -log(scores+1e-9)))."""
    d = hsc_dists(z)
    s = (F32(1) - np.exp(-d)).astype(F32)
    with np.errstate(divide="ignore", invalid="ignore"):
        anom = (-np.log(s + F32(1e-9))).astype(F32)
    losses = np.where(np.asarray(labels) == nominal_label, d, anom).astype(F32)
    return F32(losses.mean(dtype=np.float32))


def hsc_grad(z, labels, nominal_label=0):
    """Analytic d(loss)/dz of hsc_loss: (g_i / (n r_i)) z_ij, g_i = 1 (nominal) or
    -e^{-dist}/(score+1e-9) (anomalous); equals autograd through hsc.py:18-21."""
    z = _f32(z)
    n = z.shape[0]
    s2 = np.sum(z * z, axis=1, dtype=np.float32)
    r = np.sqrt(s2 + F32(1)).astype(F32)
    d = r - F32(1)
    e = np.exp(-d).astype(F32)
    sc = F32(1) - e
    with np.errstate(divide="ignore", invalid="ignore"):
        g = np.where(np.asarray(labels) == nominal_label, F32(1), -e / (sc + F32(1e-9))).astype(F32)
        coef = (g / (F32(n) * r)).astype(F32)
    return (coef[:, None] * z).astype(F32)


# --------------------------------------------------------------------------- BCE
def _sigmoid(x):
    x = _f32(x)
    out = np.empty_like(x)
    pos = x >= 0
    out[pos] = F32(1) / (F32(1) + np.exp(-x[pos]))
    ex = np.exp(x[~pos])
    out[~pos] = ex / (F32(1) + ex)
    return out.astype(F32)


def bce_score(x, nominal_label=0):
    """training/bce.py:15-17  sigmoid(f).squeeze(), or 1 - sigmoid when nominal_label != 0."""
    s = _sigmoid(np.asarray(x).reshape(-1))
    return s if nominal_label == 0 else (F32(1) - s).astype(F32)


def bce_loss(x, labels):
    """training/bce.py:19-20  binary_cross_entropy_with_logits(f.squeeze(), labels.float())
    = mean(max(x,0) - x*y + log1p(exp(-|x|))).  Labels are used raw as targets."""
    x = _f32(np.asarray(x).reshape(-1))
    y = _f32(labels)
    l = np.maximum(x, F32(0)) - x * y + np.log1p(np.exp(-np.abs(x)))
    return F32(l.astype(F32).mean(dtype=np.float32))


def bce_grad(x, labels):
    """d(loss)/dx = (sigmoid(x) - y) / n."""
    x = _f32(np.asarray(x).reshape(-1))
    y = _f32(labels)
    return ((_sigmoid(x) - y) / F32(x.shape[0])).astype(F32)


# --------------------------------------------------------------------------- DSAD
def dsad_score(z):
    """training/dsad.py:13-16: the HSC score formula."""
    return hsc_score(z)


def _dsad_dists(z):
    z = _f32(z)
    nrm = np.sqrt(np.sum(z * z, axis=1, dtype=np.float32)).astype(F32)
    return (nrm * nrm).astype(F32)                       # torch.norm(z, p=2, dim=1) ** 2   (dsad.py:19)


def dsad_loss(z, labels, nominal_label=0):
    """training/dsad.py:18-22  mean(where(labels==nominal, dists, (dists + 1e-9) ** -1))."""
    d = _dsad_dists(z)
    with np.errstate(divide="ignore"):
        anom = (F32(1) / (d + F32(1e-9))).astype(F32)
    return F32(np.where(np.asarray(labels) == nominal_label, d, anom).astype(F32).mean(dtype=np.float32))


def dsad_grad(z, labels, nominal_label=0):
    """2 z / n (nominal) or -2 z / ((dists + 1e-9)^2 n) (anomalous)."""
    z = _f32(z)
    n = z.shape[0]
    d = _dsad_dists(z)
    with np.errstate(divide="ignore", over="ignore"):
        coef = np.where(np.asarray(labels) == nominal_label, F32(2), F32(-2) / ((d + F32(1e-9)) ** 2)).astype(F32) / F32(n)
    return (coef[:, None] * z).astype(F32)


# --------------------------------------------------------------------------- DSVDD
def dsvdd_score(z, center):
    """training/dsvdd.py:23-24  (features - center).pow(2).sum(-1)."""
    diff = _f32(z) - _f32(center).reshape(1, -1)
    return np.sum(diff * diff, axis=1, dtype=np.float32).astype(F32)


def dsvdd_loss(z, center):
    """training/dsvdd.py:26-27  mean of the scores."""
    return F32(dsvdd_score(z, center).mean(dtype=np.float32))


def dsvdd_grad(z, center):
    z = _f32(z)
    return (F32(2) * (z - _f32(center).reshape(1, -1)) / F32(z.shape[0])).astype(F32)


# --------------------------------------------------------------------------- focal
def _bce_elem(x, y):
    return (np.maximum(x, F32(0)) - x * y + np.log1p(np.exp(-np.abs(x)))).astype(F32)


def focal_loss(x, labels, gamma=2.0, eps=1e-7):
    """training/focal.py:19-24  mean((1 - clamp(exp(-bce), eps, 1-eps)) ** gamma * bce)."""
    x = _f32(np.asarray(x).reshape(-1))
    y = _f32(labels)
    b = _bce_elem(x, y)
    pt = np.clip(np.exp(-b).astype(F32), F32(eps), F32(1) - F32(eps))
    return F32((((F32(1) - pt) ** F32(gamma)) * b).astype(F32).mean(dtype=np.float32))


def focal_grad(x, labels, gamma=2.0, eps=1e-7):
    """(sigmoid(x) - y) [ (1-pt)^g + 1{eps <= exp(-bce) <= 1-eps} g (1-pt)^(g-1) pt bce ] / n."""
    x = _f32(np.asarray(x).reshape(-1))
    y = _f32(labels)
    b = _bce_elem(x, y)
    raw = np.exp(-b).astype(F32)
    pt = np.clip(raw, F32(eps), F32(1) - F32(eps))
    inside = (raw >= F32(eps)) & (raw <= F32(1) - F32(eps))
    q = F32(1) - pt
    t = q ** F32(gamma) + np.where(inside, F32(gamma) * q ** F32(gamma - 1) * pt * b, F32(0))
    return ((_sigmoid(x) - y) * t / F32(x.shape[0])).astype(F32)


def focal_score(x, nominal_label=0):
    """training/focal.py:33-35: sigmoid, or 1 - sigmoid when nominal_label != 0."""
    return bce_score(x, nominal_label)


# --------------------------------------------------------------------------- CLIP head
def _unit_rows(a):
    a = _f32(a)
    return (a / np.sqrt(np.sum(a * a, axis=1, keepdims=True, dtype=np.float32))).astype(F32)


def clip_logits(z, center, renorm_center, scale=100.0):
    zt = _unit_rows(z)
    c = _unit_rows(center) if renorm_center else _f32(center)
    return (F32(scale) * zt) @ c.T


def clip_score(z, center, scale=100.0):
    """training/clip.py:66-79  softmax(100 * z^ @ T^.T)[:, -1]; center re-normalised (:69)."""
    lg = clip_logits(z, center, True, scale).astype(F32)
    lg = lg - lg.max(axis=1, keepdims=True)
    p = np.exp(lg)
    return (p[:, -1] / p.sum(axis=1, dtype=np.float32)).astype(F32)


def _clip_targets(lg, labels, nominal_label, loo):
    labels = np.asarray(labels)
    K = lg.shape[1]
    anom_label = 1 - nominal_label
    t = np.full(lg.shape[0], -1, dtype=np.int64)        # -1: row contributes 0 (clip.py:90,96)
    t[labels == anom_label] = K - 1
    if loo:
        nom = np.argmax(lg[:, :K - 1], axis=1)
    else:
        nom = np.zeros(lg.shape[0], dtype=np.int64)
    t[labels == nominal_label] = nom[labels == nominal_label]
    return t


def clip_oe_loss(z, labels, center, nominal_label=0, loo=False, scale=100.0):
    """training/clip.py:81-103  -mean(log_softmax(100 z^ @ center.T)[i, t_i]); center used as given
    (:82,86); t_i = K-1 (anomalous), 0 (ovr nominal) or argmax_{k<K-1} (loo nominal, :95)."""
    lg = clip_logits(z, center, False, scale).astype(F32)
    m = lg.max(axis=1, keepdims=True)
    lse = (m[:, 0] + np.log(np.exp(lg - m).sum(axis=1, dtype=np.float32))).astype(F32)
    ls = lg - lse[:, None]
    t = _clip_targets(lg, labels, nominal_label, loo)
    per = np.where(t >= 0, ls[np.arange(lg.shape[0]), np.maximum(t, 0)], F32(0)).astype(F32)
    return F32(-per.mean(dtype=np.float32))


def clip_oe_grad(z, labels, center, nominal_label=0, loo=False, scale=100.0):
    """d(loss)/dz: G=(softmax-onehot(t))/n (0 rows where t undefined); g=scale*G@c;
    dz = (g - (g.z^) z^)/||z||."""
    z = _f32(z)
    c = _f32(center)
    n = z.shape[0]
    nrm = np.sqrt(np.sum(z.astype(np.float64) ** 2, axis=1, keepdims=True))
    zt = z / nrm
    lg = scale * zt @ c.T.astype(np.float64)
    p = np.exp(lg - lg.max(axis=1, keepdims=True))
    p /= p.sum(axis=1, keepdims=True)
    t = _clip_targets(lg.astype(F32), labels, nominal_label, loo)
    G = p.copy()
    G[np.arange(n), np.maximum(t, 0)] -= 1.0
    G[t < 0] = 0.0
    G /= n
    g = scale * G @ c.astype(np.float64)
    dz = (g - (g * zt).sum(axis=1, keepdims=True) * zt) / nrm
    return dz.astype(F32)
