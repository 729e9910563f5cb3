"""Config builders shared by the tests (plain dicts with the reference's YAML keys)."""
import os
from types import SimpleNamespace

import torch
import yaml

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CONFIG_DIR = os.path.join(ROOT, "beta-vae_b200", "configs")


def ns(d):
    if isinstance(d, dict):
        return SimpleNamespace(**{k: ns(v) for k, v in d.items()})
    return d


def shipped(name):
    with open(os.path.join(CONFIG_DIR, f"{name}.yaml")) as f:
        return yaml.safe_load(f)


def small(**over):
    """Tiny path (32x32, base 8, 2 blocks, latent 8); ``section__key=value`` overrides."""
    cfg = {
        "data": {"image_size": 32, "grayscale": True},
        "model": {
            "latent_dim": 8, "base_channels": 8, "num_blocks": 2, "beta": 1.0,
            "se_reduction_ratio": 4, "encoder_norm": "layer", "activation": "relu",
            "reconstruction_loss": "mse", "use_decoder_se": True, "deterministic_overfit": False,
            "latent_reg_lambda": 0.0, "encoder_pooling": "flatten", "latent_clamp": None,
            "logvar_clamp": [-10, 5],
        },
        "loss": {"use_lpips": False, "use_ffl": False, "ffl_weight": 0.0, "ffl_alpha": 1.0, "lpips_weight": 0.0},
    }
    for k, v in over.items():
        sec, key = k.split("__")
        cfg[sec][key] = v
    return cfg


def weight_init(m):
    """What the reference's train loop applies (src/training/train.py:217-221)."""
    if isinstance(m, (torch.nn.Conv2d, torch.nn.ConvTranspose2d, torch.nn.Linear)):
        torch.nn.init.kaiming_normal_(m.weight, nonlinearity="relu")
        if m.bias is not None:
            torch.nn.init.zeros_(m.bias)
