"""Objective plug-ins with the reference's registry shape (src/eoe/training/__init__.py:8-11)."""
from .ad_trainer import ADTrainer, NanGradientsError
from .bce import BCETrainer
from .clip import ADClipTrainer
from .hsc import HSCTrainer

TRAINER = {  # maps strings to trainer classes, as `--objective` does in the reference (main/__init__.py:93-97)
    "hsc": HSCTrainer, "bce": BCETrainer, "clip": ADClipTrainer,
}
