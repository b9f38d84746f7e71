"""HSC objective behind the reference's hook API (src/eoe/training/hsc.py:6-21), computed by one fused kernel."""
import torch

from .. import ops
from .ad_trainer import ADTrainer


class HSCTrainer(ADTrainer):
    """Hypersphere classifier with outlier exposure; hooks have the reference's names, arguments and semantics."""

    def prepare_metric(self, cstr, loader, model, seed, **kwargs):
        return None                                                     # hsc.py:9-10

    def compute_anomaly_score(self, features, center, train: bool = False, **kwargs):
        cached = self._cached_scores(features)
        if cached is not None:                                          # loss() already produced them in the same kernel
            return cached
        return ops.hsc_score(features)                                  # hsc.py:12-15

    def loss(self, features, labels, center, **kwargs):
        loss, scores = ops.hsc_loss(features, labels, kwargs.get("nominal_label", 0))   # hsc.py:17-21 (+ backward)
        self._remember_scores(features, scores)
        return loss
