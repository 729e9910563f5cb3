"""Golden vectors for the train-time augmentation, produced by the UNMODIFIED reference transform
(/root/reference/src/data_processing/augmentations.py:5-18, ``get_train_transforms()`` under the shipped
configs/beta_vae_se.yaml: flip, rotation 10 degrees, brightness 0.1) running on torchvision + Pillow in the build
container.  Run there only:

    python tests/golden/make_augment_golden.py

Writes tests/golden/augment_default.npz:
    src      uint8 [N, S, S]   demo-data-like source images (already at cfg.data.image_size, so Resize is the identity)
    out      uint8 [N, S, S]   bytes of what the reference transform returned: its ToTensor floats are exactly
                               out / 255 in fp32 (asserted here), so the bytes carry everything
    seed     the torch seed set before the first sample
    flips / angles / factors   the parameters the same seed yields when drawn in torchvision's order
                               (oracle.augment_oracle.draw_params) -- stored so the kernel test needs no RNG
"""
import os
import sys

import numpy as np
import torch
from PIL import Image

REF = "/root/reference"
HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
os.environ["CONFIG_PATH"] = os.path.join(REF, "configs", "beta_vae_se.yaml")
sys.path.insert(0, os.path.join(REF, "src"))
from data_processing.augmentations import get_train_transforms  # noqa: E402  (the reference function)

sys.path.insert(0, ROOT)
from oracle import augment_oracle as ao  # noqa: E402

SEED, N, S = 20240, 8, 128


def demo_image(rng, kind):
    """scripts/generate_demo_data.py:27-47 in spirit: noisy canvas plus a class-dependent structure."""
    yy, xx = np.mgrid[0:S, 0:S]
    img = np.clip(rng.normal(0.25, 0.08, (S, S)), 0, 1)
    if kind == 0:
        img = img + 0.35 * (((xx - S / 2) ** 2 + (yy - S / 2) ** 2) < (S // 4) ** 2)
    elif kind == 1:
        img = img + 0.25 * ((yy % (S // 8)) < S // 16)
    elif kind == 2:
        img = img + 0.25 * ((((xx + yy) % (S // 6)) < S // 16) | (((xx - yy) % (S // 6)) < S // 16))
    else:
        img = img + rng.normal(0, 0.02, (S, S))
    return np.floor(255 * np.clip(img, 0, 1)).astype(np.uint8)


def main():
    rng = np.random.default_rng(5)
    src = np.stack([demo_image(rng, i % 4) for i in range(N)])
    t = get_train_transforms()
    torch.manual_seed(SEED)
    outs = [t(Image.fromarray(im, mode="L")) for im in src]          # float [1, S, S] each
    out_f32 = torch.stack(outs)[:, 0].numpy()
    out = np.round(out_f32 * 255.0).astype(np.uint8)
    assert np.array_equal(out.astype(np.float32) / np.float32(255.0), out_f32)
    torch.manual_seed(SEED)
    flips, angles, factors = ao.draw_params(N, True, 10.0, 0.1)
    redo = np.stack([ao.augment(im, f, a, b) for im, f, a, b in zip(src, flips, angles, factors)])
    assert np.array_equal(redo, out), "oracle restatement disagrees with the reference transform"
    np.savez_compressed(os.path.join(HERE, "augment_default.npz"), src=src, out=out, seed=np.int64(SEED),
                        flips=np.array(flips, dtype=np.uint8), angles=np.array(angles, dtype=np.float64),
                        factors=np.array(factors, dtype=np.float64))
    print("wrote augment_default.npz:", N, "images,", sum(flips), "flipped, angles", [round(a, 3) for a in angles[:4]])


if __name__ == "__main__":
    main()
