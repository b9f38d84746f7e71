"""Deep SVDD objective behind the reference's hook API (src/eoe/training/dsvdd.py:7-27), one fused kernel."""
import torch

from .. import ops
from .ad_trainer import ADTrainer


class DSVDDTrainer(ADTrainer):
    def prepare_metric(self, cstr, loader, model, seed, **kwargs):
        """dsvdd.py:11-21: mean over batches of the mean feature of the nominal (label 0) samples, entries with
        |c| < eps pushed to +-eps.  Per-batch means stay on the device (the reference moves them to the host)."""
        eps = kwargs.get("eps", 1e-1)
        center = []
        for imgs, lbls, _ in loader:
            imgs = imgs.to(self.device)
            with torch.no_grad():
                image_features = model(imgs[lbls.to(imgs.device) == 0])
            center.append(image_features.float().mean(0).unsqueeze(0))
        center = torch.cat(center).mean(0).unsqueeze(0).to(self.device)
        center[(abs(center) < eps) & (center < 0)] = -eps
        center[(abs(center) < eps) & (center > 0)] = eps
        return center

    def compute_anomaly_score(self, features, center, train: bool = False, **kwargs):
        cached = self._cached_scores(features)
        if cached is not None:
            return cached
        return ops.dsvdd_score(features, center)                        # dsvdd.py:23-24

    def loss(self, features, labels, center, **kwargs):
        loss, scores = ops.dsvdd_loss(features, center)                 # dsvdd.py:26-27 (+ backward)
        self._remember_scores(features, scores)
        return loss
