"""BCE objective behind the reference's hook API (src/eoe/training/bce.py:9-20), one fused kernel."""
from .. import ops
from .ad_trainer import ADTrainer


class BCETrainer(ADTrainer):
    def prepare_metric(self, cstr, loader, model, seed, **kwargs):
        return None                                                     # bce.py:12-13

    def compute_anomaly_score(self, features, center, train: bool = False, **kwargs):
        nominal_label = kwargs.get("nominal_label", 0)
        cached = self._cached_scores(features, tag=nominal_label)
        if cached is not None:
            return cached
        return ops.bce_score(features, nominal_label)                   # bce.py:15-17

    def loss(self, features, labels, center, **kwargs):
        nominal_label = kwargs.get("nominal_label", 0)
        loss, scores = ops.bce_loss(features, labels, nominal_label)    # bce.py:19-20 (+ backward)
        self._remember_scores(features, scores, tag=nominal_label)
        return loss
