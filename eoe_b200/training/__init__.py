"""Objective plug-ins with the reference's registry shape (src/eoe/training/__init__.py:1-11)."""
from .ad_trainer import ADTrainer, NanGradientsError
from .bce import BCETrainer
from .clip import ADClipTrainer
from .dsad import DSADTrainer
from .dsvdd import DSVDDTrainer
from .focal import FocalTrainer
from .hsc import HSCTrainer

TRAINER = {  # maps strings to trainer classes, as `--objective` does in the reference (main/__init__.py:93-97)
    "hsc": HSCTrainer, "bce": BCETrainer, "clip": ADClipTrainer,
    "dsvdd": DSVDDTrainer, "dsad": DSADTrainer, "focal": FocalTrainer,
}
