"""CLIPProcessor image pre-processing restated for the CPU (TEST INFRASTRUCTURE, see oracle/__init__.py).

Reference call sites: dataset_creation/finetune/embed_dataset.py:17-22, preprocessing/dataset_preprocessing.py:182-204,
dataset_creation/benchmark/benchmark_dataset.py:100-104 — all `CLIPProcessor.from_pretrained(CLIP_MODEL)(images=...)`.
The algorithm lives in third-party code that is not under /root/reference:
  * transformers==4.23.1 (env.yml:60) `CLIPFeatureExtractor.__call__` with the openai/clip-vit-large-patch14-336
    preprocessor config: convert RGB -> resize shortest edge to 336 (`int(size * long / short)` for the long edge,
    PIL BICUBIC) -> center crop 336 (`(dim - 336) // 2`) -> `astype(float32) / 255.0` -> `(x - mean) / std` in float32,
    channel first;
  * Pillow `Image.resize(..., resample=BICUBIC)` = libImaging/Resample.c: separable convolution with support scaled
    by the down-sampling factor, coefficients normalised in double and rounded to 22-bit fixed point, horizontal pass
    first, uint8 (rounded, clipped) intermediate, then the vertical pass.
`resize_bicubic_u8` restates Resample.c in integer numpy; tests pin it bit-exactly against the Pillow installed here and
against tests/golden/preprocess.npz (Pillow + the formulas above, written by oracle/make_golden.py).
"""
from __future__ import annotations

import math

import numpy as np

CLIP_MEAN = np.array([0.48145466, 0.4578275, 0.40821073])
CLIP_STD = np.array([0.26862954, 0.26130258, 0.27577711])
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
    """Resample.c precompute_coeffs + normalize_coeffs_8bpc for the full-image box: returns (ksize, bounds[out,2], kk[out,ksize])."""
    scale = filterscale = in_size / out_size
    if filterscale < 1.0:
        filterscale = 1.0
    support = 2.0 * filterscale
    ksize = int(math.ceil(support)) * 2 + 1
    bounds = np.zeros((out_size, 2), dtype=np.int64)
    kk = np.zeros((out_size, ksize), dtype=np.int64)
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
        for x in range(xmax):
            v = w[x] / ww if ww != 0.0 else w[x]
            kk[xx, x] = int(-0.5 + v * (1 << PRECISION_BITS)) if v < 0 else int(0.5 + v * (1 << PRECISION_BITS))
        bounds[xx] = (xmin, xmax)
    return ksize, bounds, kk


def _pass(img: np.ndarray, bounds, kk, out_size: int) -> np.ndarray:
    """One separable pass along axis 1 of img [rows, in, C] uint8 -> [rows, out, C] uint8 (clip8 of the rounded sum)."""
    out = np.empty((img.shape[0], out_size, img.shape[2]), dtype=np.uint8)
    src = img.astype(np.int64)
    for xx in range(out_size):
        xmin, xmax = bounds[xx]
        acc = (src[:, xmin:xmin + xmax, :] * kk[xx, :xmax, None]).sum(axis=1) + (1 << (PRECISION_BITS - 1))
        out[:, xx, :] = np.clip(acc >> PRECISION_BITS, 0, 255).astype(np.uint8)
    return out


def resize_bicubic_u8(img: np.ndarray, out_w: int, out_h: int) -> np.ndarray:
    """PIL Image.resize((out_w, out_h), BICUBIC) on an [H, W, C] uint8 array."""
    h, w = img.shape[:2]
    cur = img
    if out_w != w:
        _, bh, kh = precompute_coeffs(w, out_w)
        if out_h != h:                       # Resample.c only resamples the rows the vertical pass will read
            _, bv, _ = precompute_coeffs(h, out_h)
            first, last = int(bv[0, 0]), int(bv[-1, 0] + bv[-1, 1])
        else:
            first, last = 0, h
        part = _pass(cur[first:last], bh, kh, out_w)
        cur = np.zeros((h, out_w, img.shape[2]), dtype=np.uint8)
        cur[first:last] = part
    if out_h != h:
        _, bv, kv = precompute_coeffs(h, out_h)
        cur = _pass(cur.transpose(1, 0, 2), bv, kv, out_h).transpose(1, 0, 2)
    return np.ascontiguousarray(cur)


def resized_shape(h: int, w: int, size: int = 336):
    """transformers 4.23.1 image_utils.resize(default_to_square=False): (new_h, new_w)."""
    short, long = (w, h) if w <= h else (h, w)
    if short == size:
        return h, w
    new_short, new_long = size, int(size * long / short)
    return (new_long, new_short) if w <= h else (new_short, new_long)


def clip_preprocess(img: np.ndarray, size: int = 336, return_u8: bool = False):
    """[H, W, 3] uint8 RGB -> float32 [3, size, size] exactly as CLIPFeatureExtractor (4.23.1) produces it."""
    h, w = img.shape[:2]
    nh, nw = resized_shape(h, w, size)
    r = img if (nh, nw) == (h, w) else resize_bicubic_u8(img, nw, nh)
    top, left = (nh - size) // 2, (nw - size) // 2
    if top < 0 or left < 0:
        raise ValueError("center crop larger than the image (padding path) is not part of this pipeline")
    c = r[top:top + size, left:left + size]
    x = c.astype(np.float32) / np.float32(255.0)
    x = (x - CLIP_MEAN.astype(np.float32)) / CLIP_STD.astype(np.float32)
    x = np.ascontiguousarray(x.transpose(2, 0, 1))
    return (x, c) if return_u8 else x
