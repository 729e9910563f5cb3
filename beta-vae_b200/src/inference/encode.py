"""Embedding export on the B200 path: drop-in for the reference's ``src/inference/encode.py``
(reference: src/inference/encode.py:11-64 -- same function names, arguments, return values and
output files: ``<prefix>_mu.npy``, ``<prefix>_logvar.npy``, ``<prefix>_embeddings.csv`` with the
header ``path,label,z0..z{L-1}``).

What differs is how the device is driven.  The reference moves each batch with a blocking
``.to(device)`` and pulls ``mu`` / ``logvar`` back with ``.cpu()`` per batch, i.e. two host
synchronisations per 32 images; here the next batch's host->device copy is issued on a copy
stream while the current batch is encoded (uint8 batches are converted on the device,
``bvae_u8_to_unit_f32``), the results stay in one preallocated device buffer, and there is a single
device->host copy at the end.  ``encode`` itself is the forward-only use of the training kernels
(autograd saves nothing under ``torch.no_grad``)."""
import csv
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


def _images(batch):
    x = batch["image"]
    return torch.stack(x, 0) if isinstance(x, (list, tuple)) else x


def encode_loader(model, loader, device):
    """(Z [N, L], LV [N, L], labels, paths): posterior means / log-variances of every image of ``loader``."""
    device = torch.device(device)
    model.eval()
    labels, paths = [], []
    chunks = []                      # (mu, logvar) device tensors, one pair per batch
    copy_stream = torch.cuda.Stream(device=device)
    main = torch.cuda.current_stream(device)

    def stage(batch):
        x = _images(batch)
        if x.is_cuda:
            return x, None
        with torch.cuda.stream(copy_stream):
            xd = x.to(device, non_blocking=True)
            ev = torch.cuda.Event()
            ev.record(copy_stream)
        return xd, ev

    with torch.no_grad():
        it = iter(loader)
        cur = next(it, None)
        staged = stage(cur) if cur is not None else None
        while cur is not None:
            nxt = next(it, None)
            nxt_staged = stage(nxt) if nxt is not None else None     # in flight while `cur` is encoded
            xd, ev = staged
            if ev is not None:
                main.wait_event(ev)
                xd.record_stream(main)
            mu, logvar = model.encode(xd)
            chunks.append((mu, logvar))
            labels.extend(cur.get("label", [None] * xd.shape[0]))
            paths.extend(cur.get("path", [None] * xd.shape[0]))
            cur, staged = nxt, nxt_staged
    if not chunks:
        L = model.latent_dim
        return np.zeros((0, L), np.float32), np.zeros((0, L), np.float32), labels, paths
    Z = torch.cat([c[0] for c in chunks], dim=0).float().cpu().numpy()      # the one device->host copy
    LV = torch.cat([c[1] for c in chunks], dim=0).float().cpu().numpy()
    return Z, LV, labels, paths


def write_embeddings(Z, LV, labels, paths, prefix):
    cfg = get_config()
    out_dir = cfg.paths.tables_dir
    os.makedirs(out_dir, exist_ok=True)
    np.save(os.path.join(out_dir, f"{prefix}_mu.npy"), Z)
    np.save(os.path.join(out_dir, f"{prefix}_logvar.npy"), LV)
    csv_path = os.path.join(out_dir, f"{prefix}_embeddings.csv")
    with open(csv_path, "w", newline="") as f:
        w = csv.writer(f)
        w.writerow(["path", "label"] + [f"z{i}" for i in range(Z.shape[1])])
        for i in range(Z.shape[0]):
            w.writerow([paths[i], labels[i]] + list(Z[i]))
    return csv_path


def main():
    # the data pipeline (PIL decode, torchvision transforms) is the reference's own: put its src/ behind ours on sys.path
    from data_processing.augmentations import get_test_transforms, get_train_transforms
    from utils.brain_tumor_utils.datautils import build_dataloaders
    if not torch.cuda.is_available():
        raise RuntimeError("inference.encode (bvae_b200) needs a CUDA device; there is no CPU path")
    device = torch.device("cuda")
    train_loader, test_loader = build_dataloaders(transform_train=get_train_transforms(),
                                                  transform_test=get_test_transforms())
    model = load_model("best").to(device)
    Zt, LVt, Lt, Pt = encode_loader(model, train_loader, device)
    write_embeddings(Zt, LVt, Lt, Pt, "train_latents")
    Zv, LVv, Lv, Pv = encode_loader(model, test_loader, device)
    write_embeddings(Zv, LVv, Lv, Pv, "test_latents")


if __name__ == "__main__":
    main()
