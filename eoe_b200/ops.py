"""Functional host API over the C ABI: fused loss/score heads as torch.autograd Functions.

Every function launches exactly one hand-written sm_100a kernel from libeoe_b200.so on the current
CUDA stream; torch only allocates the output buffers.  Mirrors the math of the reference hooks
(src/eoe/training/hsc.py:12-21, bce.py:15-20, clip.py:66-103); see include/eoe_b200.h.
"""
import torch

from . import _lib as L


def _prep_features(z):
    L.require_cuda(z)
    if z.dim() != 2:
        raise L.EoeError(f"features must be [n, d], got {tuple(z.shape)}")
    if z.dtype not in L.DTYPE_CODE:
        z = z.float()
    return z.contiguous()


def _prep_labels(labels, n, device):
    if labels.device != device:
        labels = labels.to(device)
    if labels.dtype != torch.int64:
        labels = labels.long()
    labels = labels.reshape(-1).contiguous()
    if labels.numel() != n:
        raise L.EoeError(f"labels must have {n} entries, got {labels.numel()}")
    return labels


# ----------------------------------------------------------------------------------------------- HSC
def hsc_fused(z, labels, nominal_label=0, want_grad=True):
    """One kernel: (loss [], scores [n], dloss/dz [n,d] or None). No autograd bookkeeping."""
    z = _prep_features(z)
    n, d = z.shape
    labels = _prep_labels(labels, n, z.device)
    loss = torch.empty((), dtype=torch.float32, device=z.device)
    scores = torch.empty(n, dtype=torch.float32, device=z.device)
    grad = torch.empty_like(z) if want_grad else None
    ws = L.head_workspace(z.device)
    L.check(L.lib().eoe_hsc_fwd_bwd(L.ptr(z), L.dtype_code(z), L.ptr(labels), n, d, int(nominal_label),
                                    L.ptr(loss), L.ptr(scores), L.ptr(grad), L.ptr(ws), L.stream_ptr(z.device)),
            "eoe_hsc_fwd_bwd")
    return loss, scores, grad


def hsc_score(z):
    """HSCTrainer.compute_anomaly_score (hsc.py:12-15): 1 - exp(-(sqrt(||z||^2+1)-1))."""
    z = _prep_features(z.detach())
    n, d = z.shape
    scores = torch.empty(n, dtype=torch.float32, device=z.device)
    L.check(L.lib().eoe_hsc_score(L.ptr(z), L.dtype_code(z), n, d, L.ptr(scores), L.stream_ptr(z.device)),
            "eoe_hsc_score")
    return scores


class _HscLoss(torch.autograd.Function):
    @staticmethod
    def forward(ctx, z, labels, nominal_label):
        in_dtype = z.dtype
        loss, scores, grad = hsc_fused(z.detach(), labels, nominal_label, want_grad=True)
        ctx.save_for_backward(grad)
        ctx.in_dtype = in_dtype
        ctx.mark_non_differentiable(scores)
        return loss, scores

    @staticmethod
    def backward(ctx, g_loss, _g_scores):
        (grad,) = ctx.saved_tensors
        return (grad * g_loss.to(grad.dtype)).to(ctx.in_dtype), None, None


def hsc_loss(z, labels, nominal_label=0):
    """HSCTrainer.loss (hsc.py:17-21) with its backward fused in; returns (loss, scores)."""
    return _HscLoss.apply(z, labels, nominal_label)


# ----------------------------------------------------------------------------------------------- BCE
def _prep_logits(x):
    L.require_cuda(x)
    if x.dtype not in L.DTYPE_CODE:
        x = x.float()
    return x.reshape(-1).contiguous()       # features.squeeze() of [n,1] (bce.py:16,20)


def bce_fused(x, labels, nominal_label=0, want_grad=True):
    shape = x.shape
    x = _prep_logits(x)
    n = x.numel()
    labels = _prep_labels(labels, n, x.device)
    loss = torch.empty((), dtype=torch.float32, device=x.device)
    scores = torch.empty(n, dtype=torch.float32, device=x.device)
    grad = torch.empty_like(x) if want_grad else None
    ws = L.head_workspace(x.device)
    L.check(L.lib().eoe_bce_fwd_bwd(L.ptr(x), L.dtype_code(x), L.ptr(labels), n, int(nominal_label), L.ptr(loss),
                                    L.ptr(scores), L.ptr(grad), L.ptr(ws), L.stream_ptr(x.device)),
            "eoe_bce_fwd_bwd")
    return loss, scores, (grad.reshape(shape) if grad is not None else None)


def bce_score(x, nominal_label=0):
    """BCETrainer.compute_anomaly_score (bce.py:15-17)."""
    x = _prep_logits(x.detach())
    n = x.numel()
    scores = torch.empty(n, dtype=torch.float32, device=x.device)
    L.check(L.lib().eoe_bce_score(L.ptr(x), L.dtype_code(x), n, int(nominal_label), L.ptr(scores),
                                  L.stream_ptr(x.device)), "eoe_bce_score")
    return scores


class _BceLoss(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, labels, nominal_label):
        in_dtype = x.dtype
        loss, scores, grad = bce_fused(x.detach(), labels, nominal_label, want_grad=True)
        ctx.save_for_backward(grad)
        ctx.in_dtype = in_dtype
        ctx.mark_non_differentiable(scores)
        return loss, scores

    @staticmethod
    def backward(ctx, g_loss, _g_scores):
        (grad,) = ctx.saved_tensors
        return (grad * g_loss.to(grad.dtype)).to(ctx.in_dtype), None, None


def bce_loss(x, labels, nominal_label=0):
    """BCETrainer.loss (bce.py:19-20) with backward fused in; returns (loss, scores)."""
    return _BceLoss.apply(x, labels, nominal_label)


# ----------------------------------------------------------------------------------------------- CLIP
def _prep_text(center, device):
    L.require_cuda(center)
    return center.detach().to(device=device, dtype=torch.float32).contiguous()


def clip_score(z, center, scale=100.0):
    """ADClipTrainer.compute_anomaly_score (clip.py:66-79): softmax(100 z^ T^.T)[:, -1]."""
    z = _prep_features(z.detach())
    text = _prep_text(center, z.device)
    n, d = z.shape
    K = text.shape[0]
    scores = torch.empty(n, dtype=torch.float32, device=z.device)
    L.check(L.lib().eoe_clip_score(L.ptr(z), L.dtype_code(z), L.ptr(text), n, d, K, float(scale), L.ptr(scores),
                                   L.stream_ptr(z.device)), "eoe_clip_score")
    return scores


def clip_oe_fused(z, labels, center, nominal_label=0, leave_one_out=False, scale=100.0, want_grad=True):
    z = _prep_features(z)
    text = _prep_text(center, z.device)
    n, d = z.shape
    K = text.shape[0]
    labels = _prep_labels(labels, n, z.device)
    loss = torch.empty((), dtype=torch.float32, device=z.device)
    grad = torch.empty_like(z) if want_grad else None
    ws = L.head_workspace(z.device)
    L.check(L.lib().eoe_clip_oe_loss_fwd_bwd(L.ptr(z), L.dtype_code(z), L.ptr(text), L.ptr(labels), n, d, K,
                                             float(scale), int(nominal_label), int(bool(leave_one_out)),
                                             L.ptr(loss), L.ptr(grad), L.ptr(ws), L.stream_ptr(z.device)),
            "eoe_clip_oe_loss_fwd_bwd")
    return loss, grad


class _ClipOeLoss(torch.autograd.Function):
    @staticmethod
    def forward(ctx, z, labels, center, nominal_label, leave_one_out, scale):
        in_dtype = z.dtype
        loss, grad = clip_oe_fused(z.detach(), labels, center, nominal_label, leave_one_out, scale, True)
        ctx.save_for_backward(grad)
        ctx.in_dtype = in_dtype
        return loss

    @staticmethod
    def backward(ctx, g_loss):
        (grad,) = ctx.saved_tensors
        return (grad * g_loss.to(grad.dtype)).to(ctx.in_dtype), None, None, None, None, None


def clip_oe_loss(z, labels, center, nominal_label=0, leave_one_out=False, scale=100.0):
    """ADClipTrainer.loss (clip.py:81-103) with backward w.r.t. the image features fused in."""
    return _ClipOeLoss.apply(z, labels, center, nominal_label, leave_one_out, scale)


# ----------------------------------------------------------------------------------------------- DSAD
def dsad_fused(z, labels, nominal_label=0, want_grad=True):
    """One kernel: DSADTrainer.loss (dsad.py:18-22), its backward and the anomaly scores (dsad.py:13-16)."""
    z = _prep_features(z)
    n, d = z.shape
    labels = _prep_labels(labels, n, z.device)
    loss = torch.empty((), dtype=torch.float32, device=z.device)
    scores = torch.empty(n, dtype=torch.float32, device=z.device)
    grad = torch.empty_like(z) if want_grad else None
    ws = L.head_workspace(z.device)
    L.check(L.lib().eoe_dsad_fwd_bwd(L.ptr(z), L.dtype_code(z), L.ptr(labels), n, d, int(nominal_label),
                                     L.ptr(loss), L.ptr(scores), L.ptr(grad), L.ptr(ws), L.stream_ptr(z.device)),
            "eoe_dsad_fwd_bwd")
    return loss, scores, grad


class _DsadLoss(torch.autograd.Function):
    @staticmethod
    def forward(ctx, z, labels, nominal_label):
        in_dtype = z.dtype
        loss, scores, grad = dsad_fused(z.detach(), labels, nominal_label, want_grad=True)
        ctx.save_for_backward(grad)
        ctx.in_dtype = in_dtype
        ctx.mark_non_differentiable(scores)
        return loss, scores

    @staticmethod
    def backward(ctx, g_loss, _g_scores):
        (grad,) = ctx.saved_tensors
        return (grad * g_loss.to(grad.dtype)).to(ctx.in_dtype), None, None


def dsad_loss(z, labels, nominal_label=0):
    """DSADTrainer.loss with backward fused in; returns (loss, scores).  The score alone is `hsc_score` (same formula)."""
    return _DsadLoss.apply(z, labels, nominal_label)


# ----------------------------------------------------------------------------------------------- DSVDD
def _prep_center(center, d, device):
    L.require_cuda(center)
    c = center.detach().to(device=device, dtype=torch.float32).reshape(-1).contiguous()
    if c.numel() != d:
        raise L.EoeError(f"center must have {d} entries, got {c.numel()}")
    return c


def dsvdd_fused(z, center, want_loss=True, want_grad=True):
    """(loss [] or None, scores [n], dloss/dz or None) of DSVDDTrainer (dsvdd.py:23-27) in one kernel."""
    z = _prep_features(z)
    n, d = z.shape
    c = _prep_center(center, d, z.device)
    loss = torch.empty((), dtype=torch.float32, device=z.device) if want_loss else None
    scores = torch.empty(n, dtype=torch.float32, device=z.device)
    grad = torch.empty_like(z) if want_grad else None
    ws = L.head_workspace(z.device) if want_loss else None
    L.check(L.lib().eoe_dsvdd_fwd_bwd(L.ptr(z), L.dtype_code(z), L.ptr(c), n, d, L.ptr(loss), L.ptr(scores),
                                      L.ptr(grad), L.ptr(ws), L.stream_ptr(z.device)), "eoe_dsvdd_fwd_bwd")
    return loss, scores, grad


def dsvdd_score(z, center):
    """DSVDDTrainer.compute_anomaly_score (dsvdd.py:23-24): sum((z - c)^2, -1)."""
    return dsvdd_fused(z.detach(), center, want_loss=False, want_grad=False)[1]


class _DsvddLoss(torch.autograd.Function):
    @staticmethod
    def forward(ctx, z, center):
        in_dtype = z.dtype
        loss, scores, grad = dsvdd_fused(z.detach(), center, True, True)
        ctx.save_for_backward(grad)
        ctx.in_dtype = in_dtype
        ctx.mark_non_differentiable(scores)
        return loss, scores

    @staticmethod
    def backward(ctx, g_loss, _g_scores):
        (grad,) = ctx.saved_tensors
        return (grad * g_loss.to(grad.dtype)).to(ctx.in_dtype), None


def dsvdd_loss(z, center):
    """DSVDDTrainer.loss (dsvdd.py:26-27) with backward w.r.t. the features fused in; returns (loss, scores)."""
    return _DsvddLoss.apply(z, center)


# ----------------------------------------------------------------------------------------------- focal
def focal_fused(x, labels, nominal_label=0, gamma=2.0, eps=1e-7, want_grad=True):
    shape = x.shape
    x = _prep_logits(x)
    n = x.numel()
    labels = _prep_labels(labels, n, x.device)
    loss = torch.empty((), dtype=torch.float32, device=x.device)
    scores = torch.empty(n, dtype=torch.float32, device=x.device)
    grad = torch.empty_like(x) if want_grad else None
    ws = L.head_workspace(x.device)
    L.check(L.lib().eoe_focal_fwd_bwd(L.ptr(x), L.dtype_code(x), L.ptr(labels), n, int(nominal_label), float(gamma),
                                      float(eps), L.ptr(loss), L.ptr(scores), L.ptr(grad), L.ptr(ws),
                                      L.stream_ptr(x.device)), "eoe_focal_fwd_bwd")
    return loss, scores, (grad.reshape(shape) if grad is not None else None)


class _FocalLoss(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, labels, nominal_label, gamma, eps):
        in_dtype = x.dtype
        loss, scores, grad = focal_fused(x.detach(), labels, nominal_label, gamma, eps, want_grad=True)
        ctx.save_for_backward(grad)
        ctx.in_dtype = in_dtype
        ctx.mark_non_differentiable(scores)
        return loss, scores

    @staticmethod
    def backward(ctx, g_loss, _g_scores):
        (grad,) = ctx.saved_tensors
        return (grad * g_loss.to(grad.dtype)).to(ctx.in_dtype), None, None, None, None


def focal_loss(x, labels, nominal_label=0, gamma=2.0, eps=1e-7):
    """FocalTrainer.loss (focal.py:37-38, FocalLoss :11-24) with backward fused in; returns (loss, scores).
    The score alone is `bce_score` (sigmoid, focal.py:33-35)."""
    return _FocalLoss.apply(x, labels, nominal_label, gamma, eps)
