"""Generate golden vectors by running the UNMODIFIED reference (jtang25/Beta-VAE at
/root/reference) on seeded inputs.  Run in the build container only:

    python tests/golden/make_golden.py

The reference ships no tests / known-answer vectors, so these
files are what pins ``oracle/beta_vae_oracle.py`` and, through it, the CUDA path.
Outputs (committed):
    tests/golden/small_<case>.pt   full tensors for small configs (weights, x, eps,
                                   loss dict, every gradient)
    tests/golden/fullsize.json     loss scalars + per-tensor gradient norms for the shipped
                                   default / overfit_capacity YAMLs (SURVEY.md 8c recipe)
eps injection follows SURVEY.md 8c: the reference draws eps inside ``reparameterize``
(src/models/beta_vae.py:147-150), so the instance attribute is overridden.
"""
import json
import os
import sys
from types import SimpleNamespace

import torch
import yaml

REF = "/root/reference"
HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(REF, "src"))

from models.beta_vae import BetaVAE  # noqa: E402  (the reference class)


def ns(d):
    if isinstance(d, dict):
        return SimpleNamespace(**{k: ns(v) for k, v in d.items()})
    return d


def weight_init(m):  # what src/training/train.py:217-221 applies; restated, not imported,
    # because importing train.py pulls torchvision/sklearn side effects we do not need here
    if isinstance(m, (torch.nn.Conv2d, torch.nn.ConvTranspose2d, torch.nn.Linear)):
        torch.nn.init.kaiming_normal_(m.weight, nonlinearity="relu")
        if m.bias is not None:
            torch.nn.init.zeros_(m.bias)


def ffl64(pred, target, alpha, eps):
    """The reference's FocalFrequencyLoss hard-casts to float32 (losses.py:20-21); for the fp64
    run the same formula is evaluated without that cast."""
    d = torch.fft.fft2(pred, norm="ortho") - torch.fft.fft2(target, norm="ortho")
    dist = d.real.pow(2) + d.imag.pow(2)
    denom = dist.mean(dim=(0, 2, 3), keepdim=True) + eps
    return ((dist / denom).clamp_min(eps).pow(alpha) * dist).mean()


def base_cfg(**over):
    cfg = {
        "data": {"image_size": 32, "grayscale": True},
        "model": {
            "latent_dim": 8, "base_channels": 8, "num_blocks": 2, "beta": 1.0,
            "se_reduction_ratio": 4, "encoder_norm": "layer", "activation": "relu",
            "reconstruction_loss": "mse", "use_decoder_se": True, "deterministic_overfit": False,
            "latent_reg_lambda": 0.0, "encoder_pooling": "flatten", "latent_clamp": None,
            "logvar_clamp": [-10, 5],
        },
        "loss": {"use_lpips": False, "use_ffl": False, "ffl_weight": 0.0, "ffl_alpha": 1.0,
                 "lpips_weight": 0.0},
    }
    for k, v in over.items():
        sec, key = k.split("__")
        cfg[sec][key] = v
    return cfg


SMALL_CASES = {
    # name: (cfg overrides, batch, loss kwargs)
    "mse_capacity_ffl": (dict(loss__use_ffl=True, loss__ffl_weight=0.5),
                         3, dict(beta=1.0, capacity=3.0, free_bits=0.0, capacity_weight=1.0)),
    "l1_freebits_nodecse": (dict(model__num_blocks=3, model__se_reduction_ratio=16,
                                 model__use_decoder_se=False, model__reconstruction_loss="l1",
                                 loss__use_ffl=True, loss__ffl_weight=1.0),
                            4, dict(beta=0.15, capacity=None, free_bits=0.5, capacity_weight=None)),
    "bce_beta_clampz_reg": (dict(model__reconstruction_loss="bce", model__latent_clamp=1.5,
                                 model__latent_reg_lambda=0.01, model__logvar_clamp=None),
                            2, dict(beta=4.0, capacity=None, free_bits=0.0, capacity_weight=None)),
    "rgb_mse_beta": (dict(data__grayscale=False, data__image_size=16, model__latent_dim=4),
                     5, dict(beta=1.0, capacity=None, free_bits=0.0, capacity_weight=None)),
    "deterministic": (dict(model__deterministic_overfit=True),
                      2, dict(beta=1.0, capacity=None, free_bits=0.0, capacity_weight=None)),
    "wide16_b1": (dict(model__base_channels=16, model__num_blocks=3, data__image_size=64,
                       model__se_reduction_ratio=8, model__latent_dim=16),
                  1, dict(beta=2.0, capacity=None, free_bits=0.0, capacity_weight=None)),
}


def run_reference(cfg_obj, seed, batch, kw, logvar_scale=None, mode="fp32"):
    """mode: fp32 (the reference as shipped for mixed_precision: false), fp64 (same module in
    double: the exact answer both fp32 implementations approximate), bf16 (the reference under
    torch.autocast(cpu, bfloat16): how far the reference's own low-precision run is from exact)."""
    torch.manual_seed(seed)
    m = BetaVAE(cfg_obj)
    m.apply(weight_init)
    m.train()
    if logvar_scale is not None:  # push some logvars outside the clamp so its zero-grad branch is hit
        with torch.no_grad():
            m.fc_logvar.weight.mul_(logvar_scale)
            for mod in m.modules():  # and make norm affine / biases non-trivial
                if isinstance(mod, torch.nn.GroupNorm):
                    mod.weight.add_(0.1 * torch.randn_like(mod.weight))
                    mod.bias.add_(0.1 * torch.randn_like(mod.bias))
                if isinstance(mod, (torch.nn.Conv2d, torch.nn.Linear)):
                    mod.bias.add_(0.05 * torch.randn_like(mod.bias))
    g = torch.Generator().manual_seed(99)
    size = cfg_obj.data.image_size
    in_ch = 1 if cfg_obj.data.grayscale else 3
    x = torch.rand(batch, in_ch, size, size, generator=g)
    eps = torch.randn(batch, m.latent_dim, generator=g)
    m.reparameterize = lambda mu, logvar: mu + eps * torch.exp(0.5 * logvar)
    if mode == "fp64":
        m = m.double()
        x, eps = x.double(), eps.double()
        m.reparameterize = lambda mu, logvar: mu + eps * torch.exp(0.5 * logvar)
        m.ffl_loss.forward = lambda p, t: ffl64(p, t, m.ffl_loss.alpha, m.ffl_loss.eps)
        out = m.loss(x, **kw)
    elif mode == "bf16":
        with torch.autocast("cpu", dtype=torch.bfloat16):
            out = m.loss(x, **kw)
    else:
        out = m.loss(x, **kw)
    out["total"].backward()
    grads = {k: (p.grad.clone() if p.grad is not None else torch.zeros_like(p))
             for k, p in m.named_parameters()}
    sd = {k: v.detach().clone() for k, v in m.state_dict().items()}
    return m, sd, x, eps, out, grads


def main():
    torch.set_num_threads(8)
    for name, (over, batch, kw) in SMALL_CASES.items():
        cfg = base_cfg(**over)
        _, sd, x, eps, out, grads = run_reference(ns(cfg), seed=7, batch=batch, kw=kw, logvar_scale=6.0)
        keep = {}
        for k, v in out.items():
            if torch.is_tensor(v):
                keep[k] = v.detach().clone()
            else:
                keep[k] = v
        _, _, _, _, out64, grads64 = run_reference(ns(cfg), seed=7, batch=batch, kw=kw, logvar_scale=6.0, mode="fp64")
        extra = {"grads64": {k: v.float() for k, v in grads64.items()},
                 "grads64_resid": {k: (v - v.float().double()).float() for k, v in grads64.items()},
                 "total64": float(out64["total"])}
        if cfg["model"]["reconstruction_loss"] != "bce":
            _, _, _, _, outb, gradsb = run_reference(ns(cfg), seed=7, batch=batch, kw=kw, logvar_scale=6.0, mode="bf16")
            extra["grads_bf16"] = {k: v.float() for k, v in gradsb.items()}
            extra["total_bf16"] = float(outb["total"])
        torch.save({"cfg": cfg, "loss_kwargs": kw, "state": sd, "x": x, "eps": eps,
                    "out": keep, "grads": grads, **extra}, os.path.join(HERE, f"small_{name}.pt"))
        print(name, "total", float(out["total"]), "kl", float(out["kl_mean"]))

    full = {}
    for cfg_name, runs in {
        "beta_vae_se": [("capacity", dict(beta=1.0, capacity=30.0, free_bits=0.0, capacity_weight=1.0)),
                        ("beta4", dict(beta=4.0, capacity=None, free_bits=0.0, capacity_weight=None))],
        "overfit_capacity": [("beta4", dict(beta=4.0, capacity=None, free_bits=0.0, capacity_weight=None)),
                             ("capacity", dict(beta=1.0, capacity=30.0, free_bits=0.0, capacity_weight=1.0))],
    }.items():
        with open(os.path.join(REF, "configs", f"{cfg_name}.yaml")) as f:
            raw = yaml.safe_load(f)
        for tag, kw in runs:
            _, sd, x, eps, out, grads = run_reference(ns(raw), seed=1234, batch=4, kw=kw)
            gn = {k: float(v.norm()) for k, v in grads.items()}
            full[f"{cfg_name}/{tag}"] = {
                "seed": 1234, "input_seed": 99, "batch": 4, "loss_kwargs": kw,
                "total": float(out["total"]), "recon": float(out["recon"]),
                "recon_base": float(out["recon_base"]), "recon_ffl": float(out["recon_ffl"]),
                "kl_mean": float(out["kl_mean"]), "kl_effective": float(out["kl_effective"]),
                "grad_norm": float(torch.sqrt(sum(v.pow(2).sum() for v in grads.values()))),
                "grad_norms": gn,
                "mu_0_4": [float(v) for v in out["mu"][0, :4]],
                "param_checksum": float(sum(v.double().sum() for v in sd.values())),
                "param_abs_checksum": float(sum(v.double().abs().sum() for v in sd.values())),
                "n_params": int(sum(v.numel() for v in sd.values())),
            }
            print(cfg_name, tag, full[f"{cfg_name}/{tag}"]["total"], full[f"{cfg_name}/{tag}"]["grad_norm"])
    with open(os.path.join(HERE, "fullsize.json"), "w") as f:
        json.dump(full, f, indent=1)


if __name__ == "__main__":
    main()
