"""Focal-loss objective behind the reference's hook API (src/eoe/training/focal.py:11-39), one fused kernel."""
from .. import ops
from .ad_trainer import ADTrainer


class FocalTrainer(ADTrainer):
    GAMMA, EPS = 2.0, 1e-7                                              # FocalLoss defaults, focal.py:14

    def prepare_metric(self, cstr, loader, model, seed, **kwargs):
        return None                                                     # focal.py:28-29

    def compute_anomaly_score(self, features, center, train: bool = False, **kwargs):
        nominal_label = kwargs.get("nominal_label", 0)
        cached = self._cached_scores(features, tag=nominal_label)
        if cached is not None:
            return cached
        return ops.bce_score(features, nominal_label)                   # focal.py:33-35 (sigmoid / 1 - sigmoid)

    def loss(self, features, labels, center, **kwargs):
        nominal_label = kwargs.get("nominal_label", 0)
        loss, scores = ops.focal_loss(features, labels, nominal_label, self.GAMMA, self.EPS)   # focal.py:37-38
        self._remember_scores(features, scores, tag=nominal_label)
        return loss
