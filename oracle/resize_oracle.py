"""CPU ORACLE (TEST INFRASTRUCTURE ONLY) for the input pipeline's Resize(256, bicubic) + CenterCrop(224).

The reference's loader (``rajni/run.py:62-66``) is torchvision on PIL images:
``transforms.Resize(256, interpolation=3)`` -> ``transforms.CenterCrop(224)`` -> ``ToTensor`` -> ``Normalize``.
The arithmetic lives in two third-party dependencies that are not vendored in /root/reference (unpinned:
``README.md:7-9``): torchvision (output size + crop window; 0.26.0 installed here) and Pillow (the resampler;
12.2.0 installed here).  This file restates their published algorithms in numpy, integer for integer:

  * torchvision ``_compute_resized_output_size``: shorter edge -> 256, longer edge -> int(256 * long / short);
  * torchvision ``center_crop``: top = int(round((H - 224) / 2.0)), left likewise (Python's round-half-even);
  * Pillow ``libImaging/Resample.c``: ``precompute_coeffs`` (support = 2 * max(scale, 1) for the bicubic filter,
    a = -0.5; window [int(center - support + 0.5), int(center + support + 0.5)); weights normalised in double),
    ``normalize_coeffs_8bpc`` (PRECISION_BITS = 22, round half away from zero), then a HORIZONTAL pass to a uint8
    temporary (``clip8((1 << 21) + sum(pixel * k)) >> 22`` clamped to 0..255) followed by a VERTICAL pass.

Pinned by ``tests/test_oracle.py::test_resize_oracle_matches_torchvision`` against torchvision + Pillow themselves on
seeded images of many sizes (bit-exact), so the CUDA kernel can be checked against either.
"""
from __future__ import annotations

import math

import numpy as np

PRECISION_BITS = 32 - 8 - 2


def resized_size(h: int, w: int, size: int = 256):
    """torchvision.transforms.functional._compute_resized_output_size for an int ``size`` -> (new_h, new_w)."""
    short, long = (w, h) if w <= h else (h, w)
    new_short, new_long = size, int(size * long / short)
    return (new_long, new_short) if w <= h else (new_short, new_long)


def crop_origin(h: int, w: int, crop: int = 224):
    """torchvision center_crop (image at least as large as the crop) -> (top, left)."""
    return int(round((h - crop) / 2.0)), int(round((w - crop) / 2.0))


def _bicubic(x: float) -> float:
    a = -0.5
    if x < 0.0:
        x = -x
    if x < 1.0:
        return ((a + 2.0) * x - (a + 3.0)) * x * x + 1
    if x < 2.0:
        return (((x - 5) * x + 8) * x - 4) * a
    return 0.0


def coeffs(in_size: int, out_size: int, first: int, count: int):
    """Pillow precompute_coeffs + normalize_coeffs_8bpc for output positions [first, first + count).
    -> (bounds int32 [count, 2] = (xmin, n), kk int32 [count, ksize])."""
    scale = filterscale = float(in_size) / out_size
    if filterscale < 1.0:
        filterscale = 1.0
    support = 2.0 * filterscale
    ksize = int(math.ceil(support)) * 2 + 1
    bounds = np.zeros((count, 2), np.int32)
    kk = np.zeros((count, ksize), np.int32)
    ss = 1.0 / filterscale
    for i in range(count):
        xx = first + i
        center = (xx + 0.5) * scale
        xmin = int(center - support + 0.5)
        if xmin < 0:
            xmin = 0
        xmax = int(center + support + 0.5)
        if xmax > in_size:
            xmax = in_size
        n = xmax - xmin
        w = [_bicubic((x + xmin - center + 0.5) * ss) for x in range(n)]
        ww = 0.0
        for v in w:
            ww += v
        for x in range(n):
            v = w[x] / ww if ww != 0.0 else w[x]
            kk[i, x] = int(-0.5 + v * (1 << PRECISION_BITS)) if v < 0 else int(0.5 + v * (1 << PRECISION_BITS))
        bounds[i] = (xmin, n)
    return bounds, kk


def resize_center_crop(img: np.ndarray, size: int = 256, crop: int = 224) -> np.ndarray:
    """img uint8 [H, W, 3] (decoded RGB) -> uint8 [3, crop, crop], exactly what
    Resize(size, BICUBIC) -> CenterCrop(crop) -> PILToTensor produce."""
    h, w, _ = img.shape
    nh, nw = resized_size(h, w, size)
    top, left = crop_origin(nh, nw, crop)
    bh, kh = coeffs(w, nw, left, crop)          # horizontal pass: the crop's columns only
    bv, kv = coeffs(h, nh, top, crop)           # vertical pass: the crop's rows only
    r0, r1 = int(bv[0, 0]), int(bv[-1, 0] + bv[-1, 1])
    src = img.astype(np.int64)
    tmp = np.zeros((r1 - r0, crop, 3), np.uint8)
    for i in range(crop):
        x0, n = bh[i]
        acc = (src[r0:r1, x0:x0 + n, :] * kh[i, :n].astype(np.int64)[None, :, None]).sum(axis=1) + (1 << (PRECISION_BITS - 1))
        tmp[:, i, :] = np.clip(acc >> PRECISION_BITS, 0, 255).astype(np.uint8)
    t64 = tmp.astype(np.int64)
    out = np.zeros((3, crop, crop), np.uint8)
    for i in range(crop):
        y0, n = bv[i]
        acc = (t64[y0 - r0:y0 - r0 + n] * kv[i, :n].astype(np.int64)[:, None, None]).sum(axis=0) + (1 << (PRECISION_BITS - 1))
        out[:, i, :] = np.clip(acc >> PRECISION_BITS, 0, 255).astype(np.uint8).T
    return out
