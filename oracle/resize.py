"""TEST INFRASTRUCTURE ONLY -- CPU oracle for the head of CLIP's image transform on raw uint8 pixels:

    Resize(n_px, interpolation=BICUBIC) -> CenterCrop(n_px)          reference: src/eoe/models/clip_official/clip/clip.py:58-61

The arithmetic lives in two third-party dependencies of the reference that are not under /root/reference:
  * torchvision (requirement `torchvision>=0.18.1`, src/requirements.txt; 0.26.0 here): the size rules --
    `Resize(int)`: shorter side -> n_px, longer side -> int(n_px * long / short);  `CenterCrop`: offsets
    int(round((size - n_px) / 2.0)) (Python's round-half-even);
  * Pillow (`Image.resize(..., BICUBIC)`; 12.2.0 here), libImaging/Resample.c: two separable passes (horizontal first,
    only over the source rows the vertical pass needs), Keys bicubic a = -0.5 with support 2 * max(scale, 1), coefficients
    normalised in double precision and converted to 22-bit fixed point, 8-bit intermediate image, results
    clip8((2^21 + sum(pixel * k)) >> 22).
Restated here with numpy integer arithmetic; pinned bit-for-bit against Pillow + torchvision themselves in
tests/test_oracle_resize.py (both ship in the image) and through tests/golden/resize_*.npz.
"""
import math

import numpy as np

PRECISION_BITS = 32 - 8 - 2


def _bicubic(x: float) -> float:
    a = -0.5
    if x < 0.0:
        x = -x
    if x < 1.0:
        return ((a + 2.0) * x - (a + 3.0)) * x * x + 1
    if x < 2.0:
        return (((x - 5) * x + 8) * x - 4) * a
    return 0.0


def precompute_coeffs(in_size: int, out_size: int):
    """Resample.c precompute_coeffs + normalize_coeffs_8bpc for box = (0, in_size).
    Returns (ksize, bounds [out_size, 2] (first tap, tap count), kk [out_size, ksize] int32)."""
    scale = filterscale = in_size / out_size
    if filterscale < 1.0:
        filterscale = 1.0
    support = 2.0 * filterscale
    ksize = int(math.ceil(support)) * 2 + 1
    bounds = np.zeros((out_size, 2), np.int32)
    kk = np.zeros((out_size, ksize), np.int32)
    ss = 1.0 / filterscale
    for xx in range(out_size):
        center = (xx + 0.5) * scale
        xmin = int(center - support + 0.5)
        if xmin < 0:
            xmin = 0
        xmax = int(center + support + 0.5)
        if xmax > in_size:
            xmax = in_size
        xmax -= xmin
        w = [_bicubic((x + xmin - center + 0.5) * ss) for x in range(xmax)]
        ww = 0.0
        for v in w:
            ww += v
        if ww != 0.0:
            w = [v / ww for v in w]
        for x, v in enumerate(w):
            kk[xx, x] = int(-0.5 + v * (1 << PRECISION_BITS)) if v < 0 else int(0.5 + v * (1 << PRECISION_BITS))
        bounds[xx] = (xmin, xmax)
    return ksize, bounds, kk


def _clip8(acc):
    return np.clip(acc >> PRECISION_BITS, 0, 255).astype(np.uint8)       # arithmetic shift, as clip8_lookups indexes


def resize_bicubic_u8(img: np.ndarray, out_h: int, out_w: int) -> np.ndarray:
    """Pillow `Image.resize((out_w, out_h), BICUBIC)` of an [H, W, C] uint8 image."""
    img = np.ascontiguousarray(img, dtype=np.uint8)
    H, W, C = img.shape
    need_h, need_v = out_w != W, out_h != H
    cur = img
    if need_v:
        _, vb, vk = precompute_coeffs(H, out_h)
        first, last = int(vb[0, 0]), int(vb[-1, 0] + vb[-1, 1])
    if need_h:
        _, hb, hk = precompute_coeffs(W, out_w)
        rows = cur[first:last] if need_v else cur
        tmp = np.empty((rows.shape[0], out_w, C), np.uint8)
        r32 = rows.astype(np.int64)
        for xx in range(out_w):
            x0, n = int(hb[xx, 0]), int(hb[xx, 1])
            acc = (1 << (PRECISION_BITS - 1)) + np.tensordot(r32[:, x0:x0 + n, :], hk[xx, :n].astype(np.int64), axes=([1], [0]))
            tmp[:, xx, :] = _clip8(acc)
        cur = tmp
        if need_v:
            vb = vb.copy()
            vb[:, 0] -= first
    if need_v:
        out = np.empty((out_h, cur.shape[1], C), np.uint8)
        c32 = cur.astype(np.int64)
        for yy in range(out_h):
            y0, n = int(vb[yy, 0]), int(vb[yy, 1])
            acc = (1 << (PRECISION_BITS - 1)) + np.tensordot(vk[yy, :n].astype(np.int64), c32[y0:y0 + n], axes=([0], [0]))
            out[yy] = _clip8(acc)
        cur = out
    return cur


def resized_size(h: int, w: int, n_px: int):
    """torchvision Resize(int): (new_h, new_w)."""
    short, long_ = (w, h) if w <= h else (h, w)
    new_short, new_long = n_px, int(n_px * long_ / short)
    return (new_long, new_short) if w <= h else (new_short, new_long)


def center_crop_offsets(h: int, w: int, n_px: int):
    """torchvision CenterCrop: (top, left); Python round() is round-half-even."""
    return int(round((h - n_px) / 2.0)), int(round((w - n_px) / 2.0))


def clip_resize_center_crop(img: np.ndarray, n_px: int = 224) -> np.ndarray:
    """clip.py:58-61 on an [H, W, 3] uint8 image -> [n_px, n_px, 3] uint8."""
    H, W, _ = img.shape
    nh, nw = resized_size(H, W, n_px)
    r = resize_bicubic_u8(img, nh, nw)
    top, left = center_crop_offsets(nh, nw, n_px)
    return r[top:top + n_px, left:left + n_px]
