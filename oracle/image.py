"""CPU restatement of the reference's image preprocessing -- TEST INFRASTRUCTURE ONLY (see oracle/functional.py's header).

``TransformImage`` (pretorched/transforms/utils.py:34-81) is a torchvision ``Compose``: Resize(int(floor(max(input_size) /
scale))) -> CenterCrop(max(input_size)) -> ToTensor -> ToSpaceBGR -> ToRange255 -> Normalize(mean, std).  The arithmetic
lives in third-party code the reference does not pin: **Pillow** (``Image.resize`` with BILINEAR, installed 12.2.0) and
**torchvision** (installed 0.26.0).  This file restates

* Pillow's 8-bit separable resampling (libImaging/Resample.c: ``precompute_coeffs`` -- support = filter support x max(scale, 1),
  window [int(center - support + 0.5), int(center + support + 0.5)) clipped to the image, triangle weights normalised to
  sum 1 -- ``normalize_coeffs_8bpc`` -- fixed point with PRECISION_BITS = 32 - 8 - 2 = 22, round half away from zero --
  and ``ImagingResampleHorizontal_8bpc`` / ``Vertical``: accumulator initialised to 1 << 21, ``clip8(acc >> 22)``;
  horizontal pass first, each pass rounds to uint8), and
* torchvision's output-size, centre-crop, ToTensor and Normalize rules,

in NumPy integer / fp32 arithmetic.  Pinned: ``tests/test_image_cpu.py`` checks it against Pillow + torchvision themselves
(always importable: they are dependencies of the test environment), bit-exact on uint8 and fp32, and
``oracle/make_golden.py image`` stores the reference pipeline's output for ``data/cat.jpg`` (examples/imagenet_logits.py:38-43).
"""
import math

import numpy as np

PRECISION_BITS = 32 - 8 - 2


def bilinear_coeffs(in_size, out_size):
    """Pillow ``precompute_coeffs`` + ``normalize_coeffs_8bpc`` for the BILINEAR (triangle, support 1) filter over the whole
    axis.  Returns (bounds int32 [out][2] = (xmin, count), coeffs int32 [out][ksize])."""
    scale = in_size / out_size
    filterscale = max(scale, 1.0)
    support = 1.0 * filterscale
    ksize = int(math.ceil(support)) * 2 + 1
    bounds = np.zeros((out_size, 2), dtype=np.int32)
    kk = np.zeros((out_size, ksize), dtype=np.int32)
    ss = 1.0 / filterscale
    for xx in range(out_size):
        center = (xx + 0.5) * scale
        xmin = int(center - support + 0.5)
        xmin = max(xmin, 0)
        xmax = int(center + support + 0.5)
        xmax = min(xmax, in_size)
        n = xmax - xmin
        w = np.zeros(n, dtype=np.float64)
        for x in range(n):
            a = abs((x + xmin - center + 0.5) * ss)
            w[x] = 1.0 - a if a < 1.0 else 0.0
        ww = w.sum()            # Pillow accumulates in index order; the sum of <= ksize doubles is the same either way here
        ww = 0.0
        for x in range(n):
            ww += w[x]
        if ww != 0.0:
            w = w / ww
        for x in range(n):
            v = w[x] * (1 << PRECISION_BITS)
            kk[xx, x] = int(v - 0.5) if w[x] < 0 else int(v + 0.5)
        bounds[xx] = (xmin, n)
    return bounds, kk


def _resample_axis0(img, out_size):
    """8-bit resampling along axis 0 of a uint8 array [in][...]."""
    bounds, kk = bilinear_coeffs(img.shape[0], out_size)
    out = np.empty((out_size,) + img.shape[1:], dtype=np.uint8)
    src = img.astype(np.int64)
    for xx in range(out_size):
        x0, n = bounds[xx]
        acc = np.full(img.shape[1:], 1 << (PRECISION_BITS - 1), dtype=np.int64)
        for x in range(n):
            acc += src[x0 + x] * int(kk[xx, x])
        out[xx] = np.clip(acc >> PRECISION_BITS, 0, 255).astype(np.uint8)
    return out


def pil_resize_bilinear(img_u8, out_h, out_w):
    """uint8 [H][W][C] -> uint8 [out_h][out_w][C], Pillow semantics (horizontal pass, then vertical, each rounding to uint8;
    a pass whose size does not change is skipped)."""
    h, w = img_u8.shape[:2]
    out = img_u8
    if out_w != w:
        out = np.ascontiguousarray(_resample_axis0(np.ascontiguousarray(out.transpose(1, 0, 2)), out_w).transpose(1, 0, 2))
    if out_h != h:
        out = _resample_axis0(out, out_h)
    return out


def resized_size(h, w, size):
    """torchvision Resize(int) (transforms/functional.py ``_compute_resized_output_size``): the SHORTER edge becomes ``size``."""
    short, long = (w, h) if w <= h else (h, w)
    new_short, new_long = size, int(size * long / short)
    return (new_long, new_short) if w <= h else (new_short, new_long)      # (new_h, new_w)


def center_crop_offsets(h, w, crop):
    return int(round((h - crop) / 2.0)), int(round((w - crop) / 2.0))


def transform_image(img_u8, input_size=(3, 224, 224), input_space='RGB', input_range=(0, 1), mean=(0.485, 0.456, 0.406),
                    std=(0.229, 0.224, 0.225), scale=0.875):
    """TransformImage.__call__ (transforms/utils.py:53-81, defaults: preserve_aspect_ratio, centre crop, no flips) on a decoded
    RGB uint8 image [H][W][3].  Returns fp32 [3][crop][crop]."""
    crop = max(input_size)
    nh, nw = resized_size(img_u8.shape[0], img_u8.shape[1], int(math.floor(crop / scale)))
    r = pil_resize_bilinear(img_u8, nh, nw)
    top, left = center_crop_offsets(nh, nw, crop)
    c = r[top:top + crop, left:left + crop]
    t = c.transpose(2, 0, 1).astype(np.float32) / np.float32(255)                     # ToTensor
    if input_space == 'BGR':
        t = t[::-1].copy()                                                             # ToSpaceBGR (utils.py:14-20)
    if max(input_range) == 255:
        t = t * np.float32(255)                                                        # ToRange255 (utils.py:28-31)
    m = np.asarray(mean, dtype=np.float32).reshape(3, 1, 1)
    s = np.asarray(std, dtype=np.float32).reshape(3, 1, 1)
    return (t - m) / s                                                                  # Normalize
