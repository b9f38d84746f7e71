"""CPU: oracle/resize.py (Pillow's fixed-point bicubic resampler + torchvision's Resize / CenterCrop geometry, the head of
CLIP's `_transform`, clip_official/clip/clip.py:58-61) is bit-exact against Pillow + torchvision themselves (both ship in
the image, here and on the GPU box) and against the committed fixtures."""
import os

import numpy as np
import pytest

from oracle import resize as orz

SIZES = [(32, 32), (375, 500), (500, 375), (224, 224), (300, 224), (225, 500), (64, 100), (229, 229), (233, 350), (1000, 1500)]


def _img(h, w):
    rng = np.random.default_rng(h * 10007 + w)
    a = rng.integers(0, 256, (h, w, 3), dtype=np.uint8)
    a[: h // 4, : w // 4] = 255                        # saturated blocks exercise the clip8 overshoot of the negative lobes
    a[h // 2:, w // 2:] = 0
    return a


@pytest.mark.parametrize("h,w", SIZES)
def test_resize_oracle_bit_exact_vs_pillow_and_torchvision(h, w):
    Image = pytest.importorskip("PIL.Image")
    T = pytest.importorskip("torchvision.transforms")
    a = _img(h, w)
    want = np.asarray(T.Compose([T.Resize(224, interpolation=Image.BICUBIC), T.CenterCrop(224)])(Image.fromarray(a)))
    got = orz.clip_resize_center_crop(a, 224)
    assert got.shape == (224, 224, 3) and np.array_equal(got, want)


def test_resize_oracle_matches_golden(golden_dir):
    g = np.load(os.path.join(golden_dir, "resize.npz"))
    for h, w in [(32, 32), (375, 500), (233, 350)]:
        got = orz.clip_resize_center_crop(_img(h, w), 224)
        assert int(got.astype(np.int64).sum()) == int(g[f"sum_{h}x{w}"])
        assert np.array_equal(got[::37, ::41], g[f"sample_{h}x{w}"])


def test_geometry_rules():
    assert orz.resized_size(375, 500, 224) == (224, 298)
    assert orz.resized_size(500, 375, 224) == (298, 224)
    assert orz.center_crop_offsets(224, 298, 224) == (0, 37)
    assert orz.center_crop_offsets(224, 229, 224) == (0, 2)          # 2.5 rounds half to even
    assert orz.center_crop_offsets(224, 231, 224) == (0, 4)          # 3.5 -> 4
