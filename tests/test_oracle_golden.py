"""Pin the CPU oracle against vectors produced by the unmodified reference
(tests/golden/make_golden.py).  fp32 tolerance: rel <= 1e-5 per tensor."""
import glob
import json
import os
from types import SimpleNamespace

import pytest
import torch

from oracle import beta_vae_oracle as O

GOLD = os.path.join(os.path.dirname(__file__), "golden")
SMALL = sorted(glob.glob(os.path.join(GOLD, "small_*.pt")))


def ns(d):
    if isinstance(d, dict):
        return SimpleNamespace(**{k: ns(v) for k, v in d.items()})
    return d


from parity import check, global_norm, rel_err as rel


@pytest.mark.parametrize("path", SMALL, ids=[os.path.basename(p)[6:-3] for p in SMALL])
def test_small_case_matches_reference(path):
    g = torch.load(path, weights_only=False)
    spec = O.spec_from_config(ns(g["cfg"]))
    assert [k for k, _ in O.param_shapes(spec)] == list(g["state"].keys())
    for k, shp in O.param_shapes(spec):
        assert tuple(g["state"][k].shape) == tuple(shp), k
    out, grads = O.loss_and_grads(g["state"], spec, g["x"], g["eps"], **g["loss_kwargs"])
    ref = g["out"]
    for k in ("total", "recon", "recon_base", "recon_ffl", "kl_mean", "kl_effective",
              "kl_per_dim", "mu", "logvar", "z", "recon_img"):
        assert rel(out[k], ref[k]) <= 1e-5, (k, rel(out[k], ref[k]))
    assert out["mode"] == ref["mode"]
    gn = global_norm(g["grads"].values())
    for k, gr in g["grads"].items():
        check(k, grads[k], gr, 1e-5, gn, 1e-8)


def test_upsample_stencil_matches_torch():
    x = torch.randn(2, 3, 5, 7)
    want = torch.nn.functional.interpolate(x, scale_factor=2, mode="bilinear", align_corners=False)
    assert torch.allclose(O.upsample2x_bilinear(x), want, atol=1e-6)
    x1 = torch.randn(1, 1, 1, 1)
    assert torch.allclose(O.upsample2x_bilinear(x1), x1.expand(1, 1, 2, 2))


def test_group_norm_matches_torch():
    x = torch.randn(3, 6, 4, 4) * 3 + 1
    w, b = torch.randn(6), torch.randn(6)
    want = torch.nn.functional.group_norm(x, 1, w, b, 1e-5)
    assert torch.allclose(O.group_norm_1(x, w, b), want, atol=1e-5)


def test_shapes_default_and_overfit():
    d = O.PathSpec()
    assert d.enc_out_shape == (512, 8, 8) and d.flat_dim == 32768
    assert sum(int(torch.tensor(s).prod()) for _, s in O.param_shapes(d)) == 9574585
    o = O.PathSpec(image_size=256, base_channels=16, num_blocks=5, latent_dim=32,
                   se_reduction_ratio=16, use_decoder_se=False)
    assert sum(int(torch.tensor(s).prod()) for _, s in O.param_shapes(o)) == 2388960
    assert o.dec_channels == [(256, 128), (128, 64), (64, 32), (32, 16), (16, 16)]


def test_survey_scalars_are_in_fullsize_json():
    """SURVEY.md 8c lists these survey-time values; the regenerated file must agree."""
    with open(os.path.join(GOLD, "fullsize.json")) as f:
        full = json.load(f)
    c = full["beta_vae_se/capacity"]
    assert abs(c["total"] - 1468.026978) < 2e-3
    assert abs(c["kl_mean"] - 13.549507) < 1e-4
    assert abs(c["grad_norm"] - 4882.606999) < 5e-2
    assert c["n_params"] == 9574585
    o = full["overfit_capacity/beta4"]
    assert abs(o["total"] - 14146.382812) < 2e-2
