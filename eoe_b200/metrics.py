"""Device ROC-AUC / PRC, bit-exact with the scikit-learn calls of the reference trainer
(src/eoe/training/ad_trainer.py:453-455 and :516-522) and returned in the reference's own containers
(`ROC(tpr, fpr, ths, auc)`, `PRC(prec, rec, ths, avg_prec)`; src/eoe/utils/logger.py:36-91).
"""
from typing import Optional, Tuple

import numpy as np
import torch

from . import _lib as L


class ROC(object):
    """Same fields / accessors as eoe.utils.logger.ROC (logger.py:36-62)."""

    def __init__(self, tpr, fpr, ths, auc, std: float = 0, n: int = 1):
        self.tpr, self.fpr, self.ths, self.auc, self.std, self.n = tpr, fpr, ths, auc, std, n

    def get_x(self):
        return self.fpr

    def get_y(self):
        return self.tpr

    def get_score(self):
        return self.auc


class PRC(object):
    """Same fields / accessors as eoe.utils.logger.PRC (logger.py:65-91)."""

    def __init__(self, prec, rec, ths, avg_prec, std: float = 0, n: int = 1):
        self.prec, self.rec, self.ths, self.avg_prec, self.std, self.n = prec, rec, ths, avg_prec, std, n

    def get_x(self):
        return self.rec

    def get_y(self):
        return self.prec

    def get_score(self):
        return self.avg_prec


class AucWorkspace:
    """Reusable device buffers for `roc_auc_device` (grow-only)."""

    def __init__(self):
        self.n = 0
        self.ws = None
        self.out = None
        self.info = None

    def ensure(self, n, device):
        if self.ws is None or n > self.n or self.ws.device != device:
            nbytes = L.lib().eoe_auc_workspace_bytes(n)
            self.ws = torch.empty(nbytes, dtype=torch.uint8, device=device)
            self.out = torch.empty(2, dtype=torch.float64, device=device)
            self.info = torch.empty(8, dtype=torch.int64, device=device)
            self.n = n
        return self


_default_ws = AucWorkspace()


def roc_auc_device(scores: torch.Tensor, labels: torch.Tensor, *, ignore_negative_labels: bool = False,
                   with_prc: bool = False, curves: bool = False, workspace: Optional[AucWorkspace] = None,
                   force_tiled: bool = False, force_single_cta: bool = False, force_cluster: bool = False):
    """Launch the AUC pipeline asynchronously on the current stream.

    Returns (out, info, arrays): `out` float64[2] device = (auc, average precision), `info` int64[8] device
    (n kept, n pos, n distinct, n ROC points, status bits), `arrays` dict of device curve buffers or None.
    No host synchronisation happens here.  Up to EOE_AUC_SINGLE_LAUNCH_MAX scores this is ONE kernel launch (one CTA up
    to 12 288 scores, a cluster of 8 CTAs above; `force_single_cta` / `force_cluster` / `force_tiled` select a path
    explicitly: tests compare the three bit for bit)."""
    L.require_cuda(scores, labels)
    scores = scores.detach().reshape(-1).contiguous()
    if scores.dtype not in L.DTYPE_CODE:
        scores = scores.float()
    n = scores.numel()
    labels = labels.detach().reshape(-1)
    if labels.dtype != torch.int64:
        labels = labels.long()
    labels = labels.contiguous()
    if labels.numel() != n:
        raise L.EoeError("scores and labels differ in length")
    w = (workspace or _default_ws).ensure(n, scores.device)
    arrays = None
    fpr = tpr = thr = prec = rec = pthr = None
    if curves:
        fpr = torch.empty(n + 1, dtype=torch.float64, device=scores.device)
        tpr = torch.empty(n + 1, dtype=torch.float64, device=scores.device)
        thr = torch.empty(n + 1, dtype=torch.float32, device=scores.device)
        arrays = dict(fpr=fpr, tpr=tpr, thr=thr)
        if with_prc:
            prec = torch.empty(n + 1, dtype=torch.float64, device=scores.device)
            rec = torch.empty(n + 1, dtype=torch.float64, device=scores.device)
            pthr = torch.empty(n, dtype=torch.float32, device=scores.device)
            arrays.update(prec=prec, rec=rec, pthr=pthr)
    flags = ((L.EOE_AUC_IGNORE_NEGATIVE_LABELS if ignore_negative_labels else 0) | (L.EOE_AUC_WITH_PRC if with_prc else 0)
             | (L.EOE_AUC_FORCE_TILED if force_tiled else 0) | (L.EOE_AUC_FORCE_SINGLE_CTA if force_single_cta else 0)
             | (L.EOE_AUC_FORCE_CLUSTER if force_cluster else 0))
    L.check(L.lib().eoe_auc(L.ptr(scores), L.dtype_code(scores), L.ptr(labels), n, flags, L.ptr(w.ws),
                            w.ws.numel(), L.ptr(w.out), L.ptr(w.info), L.ptr(fpr), L.ptr(tpr), L.ptr(thr),
                            L.ptr(prec), L.ptr(rec), L.ptr(pthr), L.stream_ptr(scores.device)), "eoe_auc")
    return w.out, w.info, arrays


def _raise_on_status(status: int):
    if status & L.EOE_AUC_STATUS_NONFINITE:
        raise ValueError("Input contains NaN or infinity.")      # what sklearn's check_array raises
    if status & 0x100:
        raise L.EoeError("eoe_auc: internal look-back protocol timeout")


def roc_auc(scores, labels, **kw) -> float:
    """auc(*roc_curve(labels, scores)[:2]) as a python float (one D2H of 80 bytes). NaN if single-class."""
    out, info, _ = roc_auc_device(scores, labels, **kw)
    host = torch.cat([out.view(torch.int64), info]).cpu()
    _raise_on_status(int(host[6]))
    return float(host[:2].view(torch.float64)[0])


def roc_curve_auc(scores, labels, with_prc: bool = False, ignore_negative_labels: bool = False, force_tiled: bool = False,
                  force_single_cta: bool = False, force_cluster: bool = False) -> Tuple[Optional[ROC], Optional[PRC]]:
    """What eval_cls keeps (ad_trainer.py:516-527): (ROC, PRC) or (None, None) if a class is missing."""
    out, info, arr = roc_auc_device(scores, labels, with_prc=with_prc, curves=True,
                                    ignore_negative_labels=ignore_negative_labels, force_tiled=force_tiled,
                                    force_single_cta=force_single_cta, force_cluster=force_cluster)
    info_h = info.cpu()
    _raise_on_status(int(info_h[4]))
    if int(info_h[4]) & L.EOE_AUC_STATUS_SINGLE_CLASS:
        return None, None
    out_h = out.cpu()
    npts, m = int(info_h[3]), int(info_h[2])
    score_np_dtype = {torch.float16: np.float16, torch.bfloat16: np.float32}.get(scores.dtype, np.float32)
    roc = ROC(arr["tpr"][:npts].cpu().numpy(), arr["fpr"][:npts].cpu().numpy(),
              arr["thr"][:npts].cpu().numpy().astype(score_np_dtype), float(out_h[0]))
    prc = None
    if with_prc:
        # thresholds of precision_recall_curve: the distinct scores in increasing order (the reference's logger averages
        # them over seeds / classes, utils/logger.py:103-111, so they must be an array like sklearn's)
        prc = PRC(arr["prec"][:m + 1].cpu().numpy(), arr["rec"][:m + 1].cpu().numpy(),
                  arr["pthr"][:m].cpu().numpy().astype(score_np_dtype), float(out_h[1]))
    return roc, prc
