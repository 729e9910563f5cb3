"""Sampling, latent edits and interpolation on the B200 path: drop-in for the reference's
``src/inference/generate.py`` (reference: src/inference/generate.py:18-64 -- same function names,
arguments and output files ``samples.png``, ``edit_dim{d}.png``, ``interpolation.png``).

The reference decodes one latent per Python iteration (``steps`` batch-1 decodes, :38-47 and
:49-63); a batch-1 decode leaves 147 of the 148 SMs idle, so here every sequence is built as one
``[steps, L]`` latent matrix and decoded in ONE call -- the images are identical because every
layer of the decoder is per-sample (GroupNorm(1, C) and SE pool per sample).  ``edit_latents`` /
``interpolation_latents`` expose the latent matrices so the parity tests can check them against
the reference's loop."""
import os

import numpy as np
import torch

from models.beta_vae import BetaVAE
from utils.brain_tumor_utils.config_parser import get_config
from utils.brain_tumor_utils.io import load_sharded_checkpoint


def load_model(weights="best"):
    cfg = get_config()
    path = f"{cfg.paths.models_dir}/{cfg.paths.run_id}_{weights}.pt"
    if not os.path.exists(path):     # as the reference: checkpoints are written as *_shard{i}.pt, so this falls to `latest`
        path = f"{cfg.paths.models_dir}/{cfg.paths.run_id}_latest.pt"
    payload = load_sharded_checkpoint(path, map_location="cpu")
    model = BetaVAE()
    model.load_state_dict(payload["model_state"])
    return model


def _save_grid(imgs, path, nrow):
    from torchvision.utils import save_image          # host-side PNG writer, as in the reference
    save_image(imgs.float().cpu(), path, nrow=nrow, normalize=True)


def sample_random(model, n, out_dir, seed=None, filename="samples.png"):
    if seed is not None:
        torch.manual_seed(seed)
        np.random.seed(seed)
    with torch.no_grad():
        imgs = model.sample_prior(n)
        _save_grid(imgs, os.path.join(out_dir, filename), nrow=int(np.sqrt(n)))
    return imgs


def edit_latents(base, dim, steps, span):
    """[steps, L]: ``base`` (one latent, [1, L]) with coordinate ``dim`` swept over linspace(-span, span, steps)."""
    z = base[:1].repeat(steps, 1)
    z[:, dim] = torch.linspace(-span, span, steps, device=base.device)
    return z


def interpolation_latents(z0, z1, steps):
    """[steps, L]: (1 - a) * z0 + a * z1 for a in linspace(0, 1, steps)."""
    a = torch.linspace(0, 1, steps, device=z0.device).view(steps, 1)
    return (1 - a) * z0[:1] + a * z1[:1]


def edit_tumor_factor(model, batch, device, dim, steps, span, out_dir):
    with torch.no_grad():
        mu, _ = model.encode(batch.to(device))
        grid = model.decode(edit_latents(mu[:1], dim, steps, span))        # all steps in one decode
        _save_grid(grid, os.path.join(out_dir, f"edit_dim{dim}.png"), nrow=steps)
    return grid


def interpolate(model, img_a, img_b, device, steps, out_dir):
    with torch.no_grad():
        mu, _ = model.encode(torch.cat([img_a, img_b], dim=0).to(device))
        grid = model.decode(interpolation_latents(mu[0:1], mu[1:2], steps))
        _save_grid(grid, os.path.join(out_dir, "interpolation.png"), nrow=steps)
    return grid


def main():
    import argparse
    parser = argparse.ArgumentParser(description="Generate samples/traversals from a trained Beta-VAE.")
    parser.add_argument("--config", type=str, default=None, help="Path to YAML config")
    parser.add_argument("--weights", type=str, default="best", help="Checkpoint tag (best or latest)")
    parser.add_argument("--num-samples", type=int, default=None, help="Number of prior samples to generate")
    parser.add_argument("--seed", type=int, default=None, help="Seed for sampling latent codes")
    args = parser.parse_args()
    if args.config:
        os.environ["CONFIG_PATH"] = args.config
    # the data pipeline is the reference's own: put its src/ behind ours on sys.path
    from data_processing.augmentations import get_test_transforms
    from utils.brain_tumor_utils.datautils import build_dataloaders
    if not torch.cuda.is_available():
        raise RuntimeError("inference.generate (bvae_b200) needs a CUDA device; there is no CPU path")
    cfg = get_config()
    device = torch.device("cuda")
    test_tf = get_test_transforms()
    _, test_loader = build_dataloaders(transform_train=test_tf, transform_test=test_tf)
    model = load_model(args.weights).to(device).eval()
    out_dir = cfg.paths.figures_dir
    os.makedirs(out_dir, exist_ok=True)
    n = args.num_samples or cfg.inference.sample_grid_size
    sample_random(model, n, out_dir, seed=args.seed, filename="samples.png")
    tumor_dim = cfg.inference.tumor_latent_index
    if tumor_dim is not None:
        for batch in test_loader:
            edit_tumor_factor(model, batch["image"], device, tumor_dim, steps=cfg.evaluation.traversal_steps, span=3.0,
                              out_dir=out_dir)
            break
    it = iter(test_loader)
    try:
        b1 = next(it)["image"]
        b2 = next(it)["image"]
        interpolate(model, b1[:1], b2[:1], device, steps=cfg.evaluation.traversal_steps, out_dir=out_dir)
    except StopIteration:
        pass


if __name__ == "__main__":
    main()
