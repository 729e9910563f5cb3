#!/usr/bin/env python
"""A few graphed train steps of configs/overfit_capacity.yaml (256x256, base 16, 5 blocks) for an ncu launch list
(dev tool):  ncu --metrics gpu__time_duration.sum ... python tools/profile_overfit.py [batch]"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
PKG = os.path.join(ROOT, "beta-vae_b200")
for p in (ROOT, os.path.join(PKG, "src"), PKG):
    sys.path.insert(0, p)
import torch  # noqa: E402
import yaml  # noqa: E402

import __graft_entry__ as entry  # noqa: E402


def main():
    entry.build()
    from bvae_rt.data import synthetic_batch
    from bvae_rt.engine import GraphedTrainStep, TrainEngine
    from models.beta_vae import BetaVAE
    from utils.brain_tumor_utils.config_parser import freeze
    from tools.bench_configs import weight_init
    dev = torch.device("cuda")
    with open(os.path.join(PKG, "configs", "overfit_capacity.yaml")) as f:
        cfg = yaml.safe_load(f)
    batch = int(sys.argv[1]) if len(sys.argv) > 1 else cfg["training"]["batch_size"]
    size = cfg["data"]["image_size"]
    torch.manual_seed(cfg["data"]["seed"])
    model = BetaVAE(freeze(cfg)).apply(weight_init).to(dev).train()
    eng = TrainEngine(model, lr=cfg["optimization"]["lr"], clip=cfg["training"]["grad_clip"])
    step = GraphedTrainStep(eng, (batch, 1, size, size))
    x = synthetic_batch(batch, size, seed=0, device=dev)
    kw = dict(beta=4.0, capacity=None, free_bits=0.0, capacity_weight=None)
    for _ in range(6):
        step(x, **kw)
    torch.cuda.synchronize()
    eng.close()


if __name__ == "__main__":
    main()
