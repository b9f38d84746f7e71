"""Deep SAD objective behind the reference's hook API (src/eoe/training/dsad.py:6-22), one fused kernel."""
from .. import ops
from .ad_trainer import ADTrainer


class DSADTrainer(ADTrainer):
    def prepare_metric(self, cstr, loader, model, seed, **kwargs):
        return None                                                     # dsad.py:10-11

    def compute_anomaly_score(self, features, center, train: bool = False, **kwargs):
        cached = self._cached_scores(features)
        if cached is not None:
            return cached
        return ops.hsc_score(features)                                  # dsad.py:13-16 (same formula as hsc.py:12-15)

    def loss(self, features, labels, center, **kwargs):
        loss, scores = ops.dsad_loss(features, labels, kwargs.get("nominal_label", 0))   # dsad.py:18-22 (+ backward)
        self._remember_scores(features, scores)
        return loss
