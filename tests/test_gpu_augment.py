"""Device augmentation kernel (csrc/augment.cu through the C ABI) against the numpy oracle and the reference-generated
golden bytes -- bit-exact, this is byte work -- and its place in the input pipeline (prefetch, train loop).
(reference: src/data_processing/augmentations.py:5-18)"""
import copy
import os

import numpy as np
import pytest
import torch

import cfgs
from oracle import augment_oracle as ao

pytestmark = pytest.mark.gpu
GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "augment_default.npz")


def test_kernel_reproduces_the_reference_transform_bytes():
    from bvae_rt.augment import DeviceAugment
    g = np.load(GOLD)
    x = torch.from_numpy(g["src"]).unsqueeze(1).cuda()
    params = ([bool(f) for f in g["flips"]], [float(a) for a in g["angles"]], [float(f) for f in g["factors"]])
    aug = DeviceAugment(True, 10, 0.1)
    out = aug(x, params=params)
    assert torch.equal(out.cpu()[:, 0], torch.from_numpy(g["out"]))
    unit = torch.empty(x.shape, dtype=torch.float32, device="cuda")
    aug(x, params=params, unit_out=unit)
    assert torch.equal(unit.cpu()[:, 0], torch.from_numpy(g["out"]).float() / 255.0)     # ToTensor's floats
    # drawn under the golden seed, the product reproduces the whole reference pipeline
    torch.manual_seed(int(g["seed"]))
    assert torch.equal(aug(x).cpu()[:, 0], torch.from_numpy(g["out"]))


@pytest.mark.parametrize("shape", [(5, 1, 32, 32), (3, 3, 64, 64), (2, 1, 96, 128), (4, 1, 256, 256), (1, 1, 8, 4)])
def test_kernel_matches_oracle_on_random_and_special_parameters(shape):
    from bvae_rt.augment import DeviceAugment
    B, C, H, W = shape
    rng = np.random.default_rng(B * 1000 + H)
    x = rng.integers(0, 256, shape, dtype=np.uint8)
    for rnd in range(4):
        flips = [bool(v) for v in rng.integers(0, 2, B)]
        angles = [float(a) for a in rng.uniform(-25, 25, B)]
        factors = [float(f) for f in rng.uniform(0.5, 1.6, B)]
        if rnd == 1:                                  # Pillow's special cases: copies, exact transposes, blend edges
            angles = ([0.0, 180.0, 90.0, 270.0, 360.0] * B)[:B]
            factors = ([1.0, 0.0, 1.0000001, 2.5, None] * B)[:B]
        if rnd == 2:
            factors = [None] * B
        out = DeviceAugment()(torch.from_numpy(x).cuda(), params=(flips, angles, factors)).cpu().numpy()
        for b in range(B):
            for c in range(C):
                want = ao.augment(x[b, c], flips[b], angles[b], factors[b])
                assert np.array_equal(out[b, c], want), (shape, rnd, b, c, flips[b], angles[b], factors[b])


def test_identity_parameters_leave_the_bytes_alone_and_errors_are_loud():
    from bvae_rt import lib
    from bvae_rt.augment import DeviceAugment, apply, pack_params
    x = torch.randint(0, 256, (3, 1, 64, 64), dtype=torch.uint8, device="cuda")
    out = DeviceAugment(False, 0, 0)(x)
    assert torch.equal(out, x) and out.data_ptr() != x.data_ptr()
    words = pack_params([False] * 3, [0.0] * 3, [None] * 3, 64, 64).cuda()
    with pytest.raises(lib.BvaeError):
        apply(x, words, out=x)                                    # a gather cannot run in place
    with pytest.raises(lib.BvaeError):
        apply(torch.zeros(1, 1, 8, 6, dtype=torch.uint8, device="cuda"), words[:1])      # W % 4
    with pytest.raises(ValueError):
        DeviceAugment()(x.float())


def test_prefetch_augments_behind_the_copy_and_the_train_loop_uses_it(tmp_path):
    from bvae_rt.augment import DeviceAugment
    from training.train import train
    from utils.brain_tumor_utils.config_parser import freeze
    raw = copy.deepcopy(cfgs.shipped("beta_vae_se_debug"))
    for k in ("figures_dir", "models_dir", "outputs_dir", "tables_dir"):
        raw["paths"][k] = str(tmp_path / raw["paths"][k])
    raw["model"]["base_channels"] = 16
    raw["debug"]["max_train_batches"] = 10                         # train on all four batches, the ragged tail included
    raw["augmentation"] = dict(use_augmentations=True, horizontal_flip=True, rotation_degrees=10, brightness=0.1)
    cfg = freeze(raw)
    g = torch.Generator().manual_seed(4)
    batches = [{"image": torch.randint(0, 256, (8, 1, 128, 128), dtype=torch.uint8, generator=g), "label": [0] * 8}
               for _ in range(3)]
    tail = {"image": torch.randint(0, 256, (3, 1, 128, 128), dtype=torch.uint8, generator=g), "label": [0] * 3}

    class Loader(list):
        pass

    # what the step consumes for batch 0 under seed `data.seed` = the oracle's augmentation of that batch
    seen = []
    import bvae_rt.augment as A
    real = A.apply

    def spy(x, words, out=None, unit_out=None):
        r = real(x, words, out=out, unit_out=unit_out)
        seen.append((x.detach().cpu().clone(), words.cpu().clone(), r.detach().cpu().clone()))
        return r

    A.apply = spy
    try:
        hist = train(cfg=cfg, train_loader=Loader(batches + [tail]), test_loader=Loader(batches[:1]))
    finally:
        A.apply = real
    assert len(hist) == 2 and all(np.isfinite(h["train_total"]) for h in hist)
    assert len(seen) == 2 * 4                                      # every train batch of both epochs, the ragged one too
    x0, words0, out0 = seen[0]
    assert torch.equal(x0, batches[0]["image"]) and not torch.equal(out0, x0)
    # replay batch 0's parameter words through the oracle
    w = words0.numpy()
    for b in range(8):
        a0, a1, a2, a3, a4, a5 = (int(v) for v in w[b, 1:7])
        src = x0[b, 0].numpy()
        if w[b, 0] & A.FLIP:
            src = src[:, ::-1]
        ys, xs = np.mgrid[0:128, 0:128].astype(np.int64)
        xin, yin = (a2 + ys * a1 + xs * a0) >> 16, (a5 + ys * a4 + xs * a3) >> 16
        ok = (xin >= 0) & (xin < 128) & (yin >= 0) & (yin < 128)
        want = np.zeros((128, 128), np.uint8)
        want[ok] = src[yin[ok], xin[ok]]
        want = ao.brightness(want, float(w[b, 7:8].view(np.float32)[0]))
        assert np.array_equal(out0[b, 0].numpy(), want), b
    assert DeviceAugment.from_config(cfg) is not None
