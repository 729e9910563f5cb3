"""Host side of the device augmentation (csrc/augment.cu): parameter draws and Pillow's rotation matrix.

The reference augments every sample inside its DataLoader workers with torchvision transforms on PIL images
(src/data_processing/augmentations.py:5-18: ``RandomHorizontalFlip`` -> ``RandomRotation(degrees)`` ->
``ColorJitter(brightness)``, then ``ToTensor``); at B200 step rates that per-sample Python cannot feed one GPU.  Here
the loader ships the sized uint8 images, the host draws three numbers per sample and the device does the pixel work
in one kernel on the copy stream -- producing the same bytes torchvision would have produced for those parameters.

What stays on the host is scalar: the draws (in torchvision's order on a torch CPU generator, so a seeded run with
``num_workers=0`` sees the same parameter stream as the reference) and the inverse affine matrix, which Pillow itself
computes in Python doubles before handing 16.16 fixed-point coefficients to its C loop (``Image.rotate``)."""
from __future__ import annotations

import math

import numpy as np
import torch

from . import ops
from .lib import call

FLIP, BRIGHT = 1 << 4, 1 << 5
_ONE = 65536


def rotation_coeffs(angle_deg: float, w: int, h: int):
    """(a0..a5): source column = (a2 + y*a1 + x*a0) >> 16, source row = (a5 + y*a4 + x*a3) >> 16, the walk Pillow's
    nearest-neighbour ``Image.rotate(angle_deg, expand=False)`` performs on a ``w`` x ``h`` image, including the exact
    transposes it substitutes for multiples of a right angle."""
    angle = angle_deg % 360.0
    if angle == 0:
        return (_ONE, 0, 0, 0, _ONE, 0)
    if angle == 180:
        return (-_ONE, 0, (w - 1) * _ONE, 0, -_ONE, (h - 1) * _ONE)
    if angle == 90 and w == h:                       # counter-clockwise quarter turn: out[y][x] = in[x][w-1-y]
        return (0, -_ONE, (w - 1) * _ONE, _ONE, 0, 0)
    if angle == 270 and w == h:                      # out[y][x] = in[h-1-x][y]
        return (0, _ONE, 0, -_ONE, 0, (h - 1) * _ONE)
    rad = -math.radians(angle)
    c, s = round(math.cos(rad), 15), round(math.sin(rad), 15)
    cx, cy = w / 2, h / 2
    m = [c, s, 0.0, -s, c, 0.0]
    m[2] = m[0] * -cx + m[1] * -cy + m[2] + cx
    m[5] = m[3] * -cx + m[4] * -cy + m[5] + cy

    def fix(v):
        v = v * 65536.0 + 0.5
        return int(v) if v >= 0.0 else int(math.floor(v))

    return (fix(m[0]), fix(m[1]), fix(m[2] + m[0] * 0.5 + m[1] * 0.5),
            fix(m[3]), fix(m[4]), fix(m[5] + m[3] * 0.5 + m[4] * 0.5))


def pack_params(flips, angles, factors, w: int, h: int) -> torch.Tensor:
    """The kernel's per-sample parameter words, int32 [n, 8] (host tensor): {flags, a0..a5, factor bits}."""
    n = len(angles)
    out = np.zeros((n, 8), dtype=np.int32)
    fac = np.ones(n, dtype=np.float32)
    for i in range(n):
        flags = FLIP if flips[i] else 0
        if factors[i] is not None:
            flags |= BRIGHT
            fac[i] = factors[i]
        out[i, 0] = flags
        out[i, 1:7] = rotation_coeffs(float(angles[i]), w, h)
    out[:, 7] = fac.view(np.int32)
    return torch.from_numpy(out)


class DeviceAugment:
    """The reference's train-time augmentation (``get_train_transforms`` minus Resize / ToTensor) for uint8 batches.

    ``draw(n)`` -> (flips, angles, factors), ``__call__(u8 [B,C,H,W] cuda)`` -> augmented uint8 batch (and, with
    ``unit_out``, ToTensor's floats).  ``reference_order=True`` draws sample by sample in torchvision's call order;
    False draws the three vectors in bulk (same distributions, different stream)."""

    def __init__(self, horizontal_flip=True, rotation_degrees=10.0, brightness=0.1, generator=None, reference_order=True):
        self.horizontal_flip = bool(horizontal_flip)
        self.rotation_degrees = float(rotation_degrees or 0.0)
        self.brightness = float(brightness or 0.0)
        if self.rotation_degrees < 0 or self.brightness < 0:
            raise ValueError("If degrees / brightness is a single number, it must be non negative.")   # torchvision's check
        self.generator = generator
        self.reference_order = reference_order

    @classmethod
    def from_config(cls, cfg, **kw):
        """None when the config switches augmentation off (augmentations.py:10)."""
        a = cfg.augmentation
        if not a.use_augmentations:
            return None
        return cls(horizontal_flip=a.horizontal_flip, rotation_degrees=a.rotation_degrees, brightness=a.brightness, **kw)

    def draw(self, n):
        g = self.generator
        d, b = self.rotation_degrees, self.brightness
        lo, hi = max(0.0, 1.0 - b), 1.0 + b
        if not self.reference_order:
            flips = (torch.rand(n, generator=g) < 0.5).tolist() if self.horizontal_flip else [False] * n
            angles = torch.empty(n).uniform_(-d, d, generator=g).tolist() if d > 0 else [0.0] * n
            factors = torch.empty(n).uniform_(lo, hi, generator=g).tolist() if b > 0 else [None] * n
            return flips, angles, factors
        flips, angles, factors = [], [], []
        for _ in range(n):
            flips.append(bool(torch.rand(1, generator=g) < 0.5) if self.horizontal_flip else False)
            angles.append(float(torch.empty(1).uniform_(-d, d, generator=g).item()) if d > 0 else 0.0)
            if b > 0:
                torch.randperm(4, generator=g)       # ColorJitter shuffles its four ops before drawing the factor
                factors.append(float(torch.empty(1).uniform_(lo, hi, generator=g)))
            else:
                factors.append(None)
        return flips, angles, factors

    def __call__(self, x, params=None, out=None, unit_out=None):
        if x.dtype != torch.uint8 or x.dim() != 4:
            raise ValueError("DeviceAugment takes a uint8 batch [B, C, H, W] (the dataset's bytes before ToTensor)")
        B, C, H, W = x.shape
        if params is None:
            params = self.draw(B)
        words = params if isinstance(params, torch.Tensor) else pack_params(*params, w=W, h=H)
        words = words.to(x.device, non_blocking=True)
        return apply(x, words, out=out, unit_out=unit_out)


def apply(x, words, out=None, unit_out=None):
    """Run the kernel: ``x`` uint8 [B,C,H,W] (cuda), ``words`` int32 [B,8] on the same device."""
    B, C, H, W = x.shape
    if out is None and unit_out is None:
        out = torch.empty_like(x)
    call("bvae_augment_u8", ops._p(x), ops._p(out), ops._p(unit_out), ops._p(words), B, C, H, W, ops._stream())
    return out if out is not None else unit_out
