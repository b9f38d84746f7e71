"""Both CLIP towers behind the reference model's surface: `encode_image`, `encode_text` and, as `ADClipTrainer.__init__` sets
it up (src/eoe/training/clip.py:32-33, `model.forward = model.encode_image`), `model(imgs) -> image features`.

    clip = ClipModel(official_state_dict, device="cuda")          # keys as in clip_official/clip/model.py:395-402
    center = trainer.prepare_metric(...)                           # tokenize -> clip.encode_text -> normalise (clip.py:59-62)
    scores = trainer.compute_anomaly_score(clip(imgs), center)

Both towers run the hand-written sm_100a kernels (encoder.ClipImageEncoder, text_encoder.ClipTextEncoder); a state_dict
without text-tower keys gives an image-only model whose `encode_text` raises."""
from typing import Dict

import torch
from torch import nn

from . import _lib as L
from .encoder import ClipImageEncoder
from .text_encoder import ClipTextEncoder


class ClipModel(nn.Module):
    def __init__(self, state_dict: Dict[str, torch.Tensor], device="cuda", operand_dtype=torch.float16, max_batch: int = 256,
                 **image_encoder_kwargs):
        super().__init__()
        self.visual = ClipImageEncoder(state_dict, device=device, operand_dtype=operand_dtype, max_batch=max_batch,
                                       **image_encoder_kwargs)
        self.text = (ClipTextEncoder(state_dict, device=device, operand_dtype=operand_dtype)
                     if "token_embedding.weight" in state_dict else None)

    def encode_image(self, imgs: torch.Tensor) -> torch.Tensor:
        return self.visual(imgs)

    def encode_text(self, tokens: torch.Tensor) -> torch.Tensor:
        if self.text is None:
            raise L.EoeError("this ClipModel was built from a state_dict without the text tower (token_embedding.weight ...)")
        return self.text(tokens)

    forward = encode_image

    def score(self, imgs, center, scale: float = 100.0, out=None):
        """Fused zero-shot path: image tower + ADClipTrainer.compute_anomaly_score (clip.py:66-79)."""
        return self.visual.score(imgs, center, scale, out)
