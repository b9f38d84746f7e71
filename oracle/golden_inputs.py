"""TEST INFRASTRUCTURE ONLY -- seeded synthetic inputs shared by oracle/make_golden.py (which
feeds them to the live reference) and by tests/ (which feed them to the oracle and the CUDA
path).  numpy Generator(PCG64) / torch CPU mt19937 streams are platform independent, so the
inputs are re-derived from seeds instead of being committed; tests/golden/*.npz carry input
checksums so that a drifted generator fails loudly.

Sizes follow SURVEY.md section 8(d): HSC batch = 128 normal || 128 OE (bases.py:591-597), rep_dim 256
(models/cnn.py:47), CLIP features 512 (clip_official/clip/model.py:217), features scaled by
0.05 so that dist stays in (0,1) and scores do not saturate.
"""
import numpy as np
import torch

VIT_WEIGHT_SEED = 1
VIT_IMAGE_SEED = 3
TEXT_WEIGHT_SEED = 6
TEXT_TOKEN_SEED = 7
AUC_CASES = ("f32_2000", "f16ties_2000", "few_distinct_5000", "tiny_9", "allties_64", "negzero_300")


def hsc_inputs(n=256, d=256, seed=11):
    rng = np.random.default_rng(seed)
    z = (0.05 * rng.standard_normal((n, d))).astype(np.float32)
    z[3] = 0.0                      # zero row: dist=0, score=0, grad=0
    z[7] *= 40.0                    # large norm
    y = np.r_[np.zeros(n // 2, np.int64), np.ones(n - n // 2, np.int64)]
    return z, y


def bce_inputs(n=256, seed=12):
    rng = np.random.default_rng(seed)
    x = (3.0 * rng.standard_normal((n, 1))).astype(np.float32)
    x[5, 0] = 40.0
    x[6, 0] = -40.0
    x[9, 0] = 0.0
    y = np.r_[np.zeros(n // 2, np.int64), np.ones(n - n // 2, np.int64)]
    return x, y


def dsvdd_inputs(n=256, d=256, seed=14):
    """features and a DSVDD centre as prepare_metric leaves it (|c| >= eps = 0.1, dsvdd.py:19-20)."""
    rng = np.random.default_rng(seed)
    z = (0.3 * rng.standard_normal((n, d))).astype(np.float32)
    c = (0.2 * rng.standard_normal((1, d))).astype(np.float32)
    c[(np.abs(c) < 0.1) & (c < 0)] = -0.1
    c[(np.abs(c) < 0.1) & (c > 0)] = 0.1
    z[4] = c[0]                      # a row sitting on the centre: score 0, grad 0
    return z, c


def clip_inputs(K, n=128, d=512, seed=13):
    rng = np.random.default_rng(seed + K)
    z = rng.standard_normal((n, d)).astype(np.float32)
    c = rng.standard_normal((K, d)).astype(np.float32)
    c /= np.linalg.norm(c, axis=1, keepdims=True)       # prepare_metric returns unit rows (clip.py:62)
    c = c.astype(np.float32)
    z[: n // 2] += 0.15 * np.sqrt(d) * c[rng.integers(0, K, n // 2)]   # make some prompts win clearly
    y = (rng.random(n) < 0.5).astype(np.int64)
    y[1] = 2                                             # label outside {0,1}: contributes 0 (clip.py:90-92)
    return z, y, c


def auc_inputs(name):
    rng = np.random.default_rng(abs(hash_name(name)))
    if name == "f32_2000":
        n = 2000
        s = (1 - np.exp(-np.abs(rng.standard_normal(n)))).astype(np.float32)
        y = (rng.random(n) < 0.3).astype(np.int64)
    elif name == "f16ties_2000":
        n = 2000
        s = (1 - np.exp(-np.abs(rng.standard_normal(n)))).astype(np.float16).astype(np.float32)
        y = (rng.random(n) < 0.5).astype(np.int64)
    elif name == "few_distinct_5000":
        n = 5000
        s = (rng.integers(0, 37, n) / 37.0).astype(np.float32)
        y = (rng.random(n) < 0.9).astype(np.int64)
    elif name == "tiny_9":
        s = np.array([0.1, 0.4, 0.35, 0.8, 0.8, 0.2, 0.9, 0.05, 0.5], np.float32)
        y = np.array([0, 0, 1, 1, 0, 0, 1, 0, 1], np.int64)
    elif name == "allties_64":
        s = np.full(64, 0.25, np.float32)
        y = (np.arange(64) % 3 == 0).astype(np.int64)
    elif name == "negzero_300":
        n = 300
        s = rng.standard_normal(n).astype(np.float32)
        s[::7] = 0.0
        s[3::7] = -0.0
        y = (rng.random(n) < 0.4).astype(np.int64)
    else:
        raise KeyError(name)
    return y, s


def hash_name(name):
    h = 0
    for ch in name.encode():
        h = (h * 131 + ch) % (2 ** 31 - 1)
    return h


def vit_images(B=2, res=224, seed=VIT_IMAGE_SEED):
    g = torch.Generator().manual_seed(seed)
    return torch.randn(B, 3, res, res, generator=g)


SCORE_PARITY_IMAGES = 64
SCORE_PARITY_CFGS = ((32, 10), (16, 30))        # BASELINE configs 2 and 3: (patch, prompts)


def score_parity_inputs(K, n_img=SCORE_PARITY_IMAGES, seed=17):
    """End-to-end score parity: seeded 224x224 images, K unit prompt rows (SURVEY 8d cfg2/cfg3:
    `normalize(randn(K,512))`), Bernoulli(0.9 anomalous) labels for the AUC of the scores."""
    imgs = vit_images(B=n_img, seed=VIT_IMAGE_SEED + 100)
    g = torch.Generator().manual_seed(seed + K)
    text = torch.nn.functional.normalize(torch.randn(K, 512, generator=g), dim=-1).numpy().astype(np.float32)
    labels = (torch.rand(n_img, generator=g) < 0.9).numpy().astype(np.int64)
    labels[0], labels[1] = 0, 1                    # both classes present whatever the draw
    return imgs, text, labels


def text_tokens(n=10, seed=TEXT_TOKEN_SEED):
    """[n, 77] prompt rows shaped like `tokenize` output (leave-one-out prompt set of cfg2: 10 prompts)."""
    from oracle import text as otext
    return otext.synth_tokens(n, seed=seed)
