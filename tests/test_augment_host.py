"""Train-time augmentation, CPU side: the numpy oracle against the reference-generated golden vectors and against
torchvision + Pillow themselves (library code in this image, not the reference tree), and the product's host logic
(bvae_rt.augment: parameter draws, Pillow's fixed-point rotation matrix) against the oracle.
(reference: src/data_processing/augmentations.py:5-18)"""
import os

import numpy as np
import pytest
import torch

from oracle import augment_oracle as ao

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "augment_default.npz")


def _gold():
    g = np.load(GOLD)
    return g, [bool(f) for f in g["flips"]], [float(a) for a in g["angles"]], [float(f) for f in g["factors"]]


def test_oracle_reproduces_the_reference_transform_bytes():
    g, flips, angles, factors = _gold()
    for i in range(len(angles)):
        assert np.array_equal(ao.augment(g["src"][i], flips[i], angles[i], factors[i]), g["out"][i]), i
    assert any(flips) and not all(flips)
    assert (g["out"] != g["src"]).mean() > 0.5                   # the transform really moved the pixels


def test_oracle_draw_order_is_torchvisions():
    g, flips, angles, factors = _gold()
    torch.manual_seed(int(g["seed"]))
    f2, a2, b2 = ao.draw_params(len(angles), True, 10.0, 0.1)
    assert f2 == flips and a2 == angles and b2 == factors


def test_product_draws_and_coefficients_match_the_oracle():
    from bvae_rt.augment import BRIGHT, FLIP, DeviceAugment, pack_params, rotation_coeffs
    g, flips, angles, factors = _gold()
    gen = torch.Generator().manual_seed(int(g["seed"]))          # an explicit generator sees the same stream
    aug = DeviceAugment(True, 10, 0.1, generator=gen)
    assert aug.draw(len(angles)) == (flips, angles, factors)
    rng = np.random.default_rng(3)
    for w, h in ((128, 128), (32, 32), (256, 128), (64, 96)):
        for ang in list(rng.uniform(-30, 30, 25)) + [0.0, 360.0, 180.0, -180.0, 90.0, 270.0, -90.0, 1e-7]:
            co = ao.affine_fixed_coeffs(float(ang), w, h)
            mine = rotation_coeffs(float(ang), w, h)
            if co is None or co[0] == "transpose":               # exact transposes: check them as pixel maps
                ys, xs = np.mgrid[0:h, 0:w]
                img = rng.integers(0, 256, (h, w), dtype=np.uint8)
                xin = (mine[2] + ys * mine[1] + xs * mine[0]) >> 16
                yin = (mine[5] + ys * mine[4] + xs * mine[3]) >> 16
                assert np.array_equal(img[yin, xin], ao.rotate_nearest(img, float(ang))), (w, h, ang)
            else:
                assert tuple(co) == tuple(mine), (w, h, ang)
    words = pack_params([True, False], [3.5, 0.0], [1.05, None], 128, 128).numpy()
    assert words.shape == (2, 8) and words.dtype == np.int32
    assert words[0, 0] == FLIP | BRIGHT and words[1, 0] == 0
    assert words[0, 7:8].view(np.float32)[0] == np.float32(1.05)
    assert tuple(words[1, 1:7]) == (65536, 0, 0, 0, 65536, 0)


def test_from_config_reads_the_reference_keys():
    import cfgs
    from bvae_rt.augment import DeviceAugment
    from utils.brain_tumor_utils.config_parser import freeze
    raw = cfgs.shipped("beta_vae_se")
    aug = DeviceAugment.from_config(freeze(raw))
    assert (aug.horizontal_flip, aug.rotation_degrees, aug.brightness) == (True, 10.0, 0.1)
    raw["augmentation"]["use_augmentations"] = False
    assert DeviceAugment.from_config(freeze(raw)) is None
    raw["augmentation"].update(use_augmentations=True, rotation_degrees=0, brightness=0, horizontal_flip=False)
    off = DeviceAugment.from_config(freeze(raw))
    assert off.draw(3) == ([False] * 3, [0.0] * 3, [None] * 3)    # nothing drawn, nothing consumed from the RNG


def test_oracle_against_live_torchvision():
    tv = pytest.importorskip("torchvision")
    Image = pytest.importorskip("PIL.Image")
    F = tv.transforms.functional
    rng = np.random.default_rng(11)
    for trial in range(60):
        s = int(rng.choice([32, 64, 128]))
        img = rng.integers(0, 256, (s, s), dtype=np.uint8)
        ang = float(rng.uniform(-10, 10)) if trial % 6 else float(rng.choice([0.0, 180.0, 90.0, 270.0, 45.0]))
        fac = float(rng.uniform(0.9, 1.1)) if trial % 5 else float(rng.choice([1.0, 0.0, 1.5, 0.5]))
        flip = bool(rng.integers(0, 2))
        p = Image.fromarray(img, mode="L")
        ref = np.asarray(F.adjust_brightness(F.rotate(F.hflip(p) if flip else p, ang), fac))
        assert np.array_equal(ao.augment(img, flip, ang, fac), ref), (trial, s, ang, flip, fac)
