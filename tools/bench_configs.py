#!/usr/bin/env python
"""Throughput of the other BASELINE.json configs (dev tool; bench.py's line is the default config at batch 32):
  * train step on configs/overfit_capacity.yaml (256x256, base 16, 5 blocks, beta = 4) at its batch 8 and at 64;
  * inference (BASELINE config 5): encode / sample_prior (decode) / full forward at batch 1024, default config, bf16.
    python tools/bench_configs.py"""
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
PKG = os.path.join(ROOT, "beta-vae_b200")
for p in (ROOT, os.path.join(PKG, "src"), PKG):
    sys.path.insert(0, p)
import torch  # noqa: E402
import yaml  # noqa: E402

import __graft_entry__ as entry  # noqa: E402


def weight_init(m):
    if isinstance(m, (torch.nn.Conv2d, torch.nn.Linear)):
        torch.nn.init.kaiming_normal_(m.weight, nonlinearity="relu")
        if m.bias is not None:
            torch.nn.init.zeros_(m.bias)


def timed(fn, steps):
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(steps):
        fn(i)
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / steps


def main():
    entry.build()
    from bvae_rt.data import synthetic_batch
    from bvae_rt.engine import GraphedTrainStep, TrainEngine
    from models.beta_vae import BetaVAE
    from utils.brain_tumor_utils.config_parser import freeze
    dev = torch.device("cuda")

    def load(name):
        with open(os.path.join(PKG, "configs", name + ".yaml")) as f:
            return yaml.safe_load(f)

    # ---- training, overfit_capacity
    cfg = load("overfit_capacity")
    size = cfg["data"]["image_size"]
    for batch in (cfg["training"]["batch_size"], 64):
        torch.manual_seed(cfg["data"]["seed"])
        model = BetaVAE(freeze(cfg)).apply(weight_init).to(dev).train()
        eng = TrainEngine(model, lr=cfg["optimization"]["lr"], clip=cfg["training"]["grad_clip"])
        step = GraphedTrainStep(eng, (batch, 1, size, size))
        xs = [synthetic_batch(batch, size, seed=i, device=dev) for i in range(4)]
        kw = dict(beta=4.0, capacity=None, free_bits=0.0, capacity_weight=None)
        for i in range(5):
            step(xs[i % 4], **kw)
        ms = timed(lambda i: step(xs[i % 4], **kw), 30)
        print(f"train overfit_capacity.yaml ({size}x{size}, base {cfg['model']['base_channels']}, bf16, graph) batch {batch:3d}: "
              f"{ms:7.3f} ms/step  {batch / ms * 1e3:9.0f} images/s", flush=True)
        eng.close()
        del model, eng, step

    # ---- inference, default config
    cfg = load("beta_vae_se")
    size, latent = cfg["data"]["image_size"], cfg["model"]["latent_dim"]
    torch.manual_seed(0)
    model = BetaVAE(freeze(cfg)).apply(weight_init).to(dev).eval()
    B = 1024
    x = synthetic_batch(B, size, seed=0, device=dev)
    with torch.no_grad(), torch.autocast("cuda", dtype=torch.bfloat16):
        for name, fn in (("encode", lambda i: model.encode(x)),
                         ("sample_prior (decode)", lambda i: model.sample_prior(B)),
                         ("forward(deterministic)", lambda i: model.forward(x, deterministic=True))):
            fn(0)
            fn(0)
            ms = timed(fn, 5)
            print(f"inference default config, bf16, batch {B}: {name:24s} {ms:8.2f} ms  {B / ms * 1e3:9.0f} images/s", flush=True)


if __name__ == "__main__":
    main()
