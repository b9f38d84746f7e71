"""CLIP zero-shot / outlier-exposure objective behind the reference's hook API (src/eoe/training/clip.py:14-103).

`model(imgs)` is the B200 image encoder (eoe_b200.encoder.ClipImageEncoder); `prepare_metric` needs a text encoder,
which the reference runs once per class (clip.py:50-64, not hot): pass `text_encoder` (a callable mapping a list of
prompts to [K, 512] features, e.g. the reference's tokenize + encode_text) or precomputed `text_features`."""
from typing import Callable, Dict, Optional, Sequence

import torch

from .. import ops
from .ad_trainer import ADTrainer


class ADClipTrainer(ADTrainer):
    def __init__(self, model, *args, anom_tkn_ptn: str = "a photo of something",
                 text_encoder: Optional[Callable[[Sequence[str]], torch.Tensor]] = None,
                 class_names: Optional[Sequence[str]] = None,
                 text_features: Optional[Dict[str, torch.Tensor]] = None, **kwargs):
        super().__init__(model, *args, **kwargs)
        self.anom_tkn_ptn = anom_tkn_ptn
        self.text_encoder = text_encoder
        self.class_names = list(class_names) if class_names is not None else None
        self.text_features = text_features or {}
        self.raw_texts = None

    def prepare_metric(self, cstr, loader, model, seed, **kwargs):
        if self.ad_mode == "one_vs_rest":                               # clip.py:51-52
            raw_texts = [f"a photo of a {cstr}", self.anom_tkn_ptn.format(cstr)]
        elif self.ad_mode == "leave_one_out":                           # clip.py:53-54
            if self.class_names is None:
                raise ValueError("leave_one_out needs class_names (str_labels(dataset) in the reference)")
            raw_texts = [*[f"a photo of a {cs}" for cs in self.class_names if cs != cstr], self.anom_tkn_ptn.format(cstr)]
        else:
            raise NotImplementedError()
        self.raw_texts = raw_texts
        if cstr in self.text_features:
            text_features = self.text_features[cstr].to(self.device).float()
        elif self.text_encoder is not None:
            with torch.no_grad():
                text_features = self.text_encoder(raw_texts).to(self.device).float()
        else:
            raise ValueError("ADClipTrainer needs `text_encoder` or precomputed `text_features[cstr]`")
        return text_features / text_features.norm(dim=-1, keepdim=True)    # clip.py:62

    def compute_anomaly_score(self, image_features, center, train: bool = False, **kwargs):
        if self.ad_mode not in self.AD_MODES:
            raise NotImplementedError()
        return ops.clip_score(image_features, center, 100.0)            # clip.py:66-79 (both modes take [:, -1])

    def loss(self, image_features, labels, center, **kwargs):
        if self.ad_mode not in self.AD_MODES:
            raise NotImplementedError()
        return ops.clip_oe_loss(image_features, labels, center, kwargs.get("nominal_label", 0),
                                leave_one_out=(self.ad_mode == "leave_one_out"), scale=100.0)   # clip.py:81-103
