"""CPU oracle for the train-time image augmentation.  TEST INFRASTRUCTURE ONLY (same rules as beta_vae_oracle.py:
imported by ``tests/`` only; the product path never touches it).

Restates, in numpy integer arithmetic, what the reference's train transform does to an already-sized 8-bit
grayscale image between ``Resize`` and ``ToTensor`` (reference: src/data_processing/augmentations.py:5-18):

    RandomHorizontalFlip()  ->  RandomRotation(degrees=d)  ->  ColorJitter(brightness=b)

The arithmetic lives in third-party code that is not under /root/reference: torchvision (unpinned in the reference's
requirements.txt; 0.26.0 in this image) drives Pillow (12.2.0 here) on ``PIL.Image`` objects of mode "L":

  * flip        ``Image.transpose(FLIP_LEFT_RIGHT)``: out[y][x] = in[y][W-1-x].
  * rotation    ``Image.rotate(angle, NEAREST, expand=False, center=None, fillcolor=0)``.  Pillow builds the inverse
                affine matrix in Python doubles (cos / sin rounded to 15 decimals, centre (w/2, h/2)), and its C core
                (Geometry.c, nearest-neighbour path) walks the destination in 16.16 FIXED POINT:
                    FIX(v) = floor(v * 65536 + 0.5);  a0..a5 = FIX of the matrix, with the half-pixel offset folded
                    into the translations:  a2 = FIX(m2 + m0/2 + m1/2),  a5 = FIX(m5 + m3/2 + m4/2);
                    source column = (a2 + y*a1 + x*a0) >> 16, source row = (a5 + y*a4 + x*a3) >> 16;
                    pixels whose source falls outside the image keep the fill colour 0.
                ``angle % 360 == 0`` returns a copy; 180 (and 90 / 270 on square images) are exact transposes.
  * brightness  ``ImageEnhance.Brightness(img).enhance(f)`` = ``Image.blend(black, img, f)``; in C (Blend.c), with
                f converted to a 32-bit float:  0 <= f <= 1: (uint8)(f * v);  otherwise clip(f * v) to [0, 255] then
                truncate.  f == 1 returns a copy.

The random draws follow torchvision's order on the global torch CPU generator, per sample:
    ``torch.rand(1) < 0.5``                          (RandomHorizontalFlip.forward)
    ``torch.empty(1).uniform_(-d, d).item()``        (RandomRotation.get_params)
    ``torch.randperm(4)``; ``torch.empty(1).uniform_(max(0, 1-b), 1+b)``   (ColorJitter.get_params)

Pinned by ``tests/golden/make_augment_golden.py`` (runs the real torchvision Compose on PIL images in the build
container and commits inputs, drawn parameters and outputs) and ``tests/test_augment_host.py``.
"""
from __future__ import annotations

import math

import numpy as np


def affine_fixed_coeffs(angle_deg: float, w: int, h: int):
    """The six 16.16 fixed-point coefficients Pillow's nearest-neighbour affine path uses for
    ``Image.rotate(angle_deg)`` on a w x h image; None when the rotation is the identity (Pillow returns a copy)."""
    angle = angle_deg % 360.0
    if angle == 0:
        return None
    if angle == 180 or (angle in (90, 270) and w == h):
        return ("transpose", angle)
    a = -math.radians(angle)
    m = [round(math.cos(a), 15), round(math.sin(a), 15), 0.0, round(-math.sin(a), 15), round(math.cos(a), 15), 0.0]
    cx, cy = w / 2, h / 2
    m[2] = m[0] * -cx + m[1] * -cy + m[2]
    m[5] = m[3] * -cx + m[4] * -cy + m[5]
    m[2] += cx
    m[5] += cy

    def fix(v):
        v = v * 65536.0 + 0.5
        return int(v) if v >= 0.0 else int(math.floor(v))

    return (fix(m[0]), fix(m[1]), fix(m[2] + m[0] * 0.5 + m[1] * 0.5), fix(m[3]), fix(m[4]),
            fix(m[5] + m[3] * 0.5 + m[4] * 0.5))


def rotate_nearest(img: np.ndarray, angle_deg: float) -> np.ndarray:
    h, w = img.shape
    co = affine_fixed_coeffs(angle_deg, w, h)
    if co is None:
        return img.copy()
    if co[0] == "transpose":
        return np.rot90(img, k=int(co[1]) // 90).copy()      # counter-clockwise, as Pillow's ROTATE_*
    a0, a1, a2, a3, a4, a5 = co
    ys, xs = np.mgrid[0:h, 0:w].astype(np.int64)
    xin = (a2 + ys * a1 + xs * a0) >> 16
    yin = (a5 + ys * a4 + xs * a3) >> 16
    ok = (xin >= 0) & (xin < w) & (yin >= 0) & (yin < h)
    out = np.zeros_like(img)
    out[ok] = img[yin[ok], xin[ok]]
    return out


def brightness(img: np.ndarray, factor: float) -> np.ndarray:
    f = np.float32(factor)
    if f == np.float32(1.0):
        return img.copy()
    if f == np.float32(0.0):
        return np.zeros_like(img)
    t = (f * img.astype(np.float32)).astype(np.float32)
    if not (np.float32(0.0) <= f <= np.float32(1.0)):
        t = np.clip(t, np.float32(0.0), np.float32(255.0))
    return t.astype(np.uint8)                                 # C cast: truncation toward zero


def augment(img: np.ndarray, flip: bool, angle_deg: float, factor) -> np.ndarray:
    """One [H, W] uint8 image through flip -> rotate -> brightness (``factor`` None: no colour jitter)."""
    out = img[:, ::-1].copy() if flip else img.copy()
    out = rotate_nearest(out, angle_deg)
    if factor is not None:
        out = brightness(out, factor)
    return out


def draw_params(n: int, horizontal_flip=True, rotation_degrees=10.0, brightness_range=0.1, generator=None):
    """Per-sample (flip, angle, factor) in torchvision's draw order on a torch CPU generator (None = the global one)."""
    import torch
    flips, angles, factors = [], [], []
    for _ in range(n):
        flip = False
        if horizontal_flip:
            flip = bool(torch.rand(1, generator=generator) < 0.5)
        ang = 0.0
        if rotation_degrees and rotation_degrees > 0:
            ang = float(torch.empty(1).uniform_(-float(rotation_degrees), float(rotation_degrees), generator=generator).item())
        fac = None
        if brightness_range and brightness_range > 0:
            torch.randperm(4, generator=generator)
            fac = float(torch.empty(1).uniform_(max(0.0, 1.0 - brightness_range), 1.0 + brightness_range,
                                                generator=generator))
        flips.append(flip)
        angles.append(ang)
        factors.append(fac)
    return flips, angles, factors
