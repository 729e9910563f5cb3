"""Training loop for the B200 path -- same semantics as the reference's ``training.train.train``
(reference: src/training/train.py:223-470) with the step itself replaced by the engine:

  reference                                             here
  ---------------------------------------------------   ----------------------------------------------
  optimizer.zero_grad / autocast(model.loss) /          ``GraphedTrainStep``: one CUDA-graph replay per
  GradScalerWrapper.backward (clip, Adam)  (:313-330)   step (bf16 kernels when training.mixed_precision)
  8 ``.item()`` host syncs per step  (:335-343)         running sums stay on the device; one read-back per
                                                        ``logging.log_every_n_steps``
  per-epoch beta / capacity / free-bits  (:296-300)     same schedulers (training/schedulers.py)
  CosineAnnealingLR per epoch, StepLR per step          closed forms of the same schedules -> engine.lr
  validation pass without autocast  (:386-413)          same: exact-fp32 kernels
  CheckpointManager latest/best, 2 shards  (:452-453)   same files / payload keys (utils/.../io.py)
  EarlyStopping(patience=20)  (:466-468)                same
  --resume best|latest  (:266-280)                      same

Multi-GPU: launch one process per GPU (torchrun); every rank runs this function on its own loader shard, gradients
are all-reduced by the engine, rank 0 logs and writes checkpoints.  Out of scope here: the PIL/torchvision dataset
(pass any iterable of ``{"image": tensor [B,C,H,W], "label": ...}`` batches -- floats in [0,1] as ToTensor leaves them,
or the sized uint8 bytes, which are augmented per cfg.augmentation and scaled on the device; without a loader,
synthetic MRI-shaped batches are generated on the device), the sklearn probe metrics and the reconstruction PNG panels."""
import argparse
import math
import os
import random
import sys

_SRC = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))          # .../beta-vae_b200/src
for _p in (_SRC, os.path.dirname(_SRC)):                                     # models/training/utils, bvae_rt
    if _p not in sys.path:
        sys.path.insert(0, _p)

import numpy as np
import torch
import torch.distributed as dist

from bvae_rt.augment import DeviceAugment
from bvae_rt.data import synthetic_batch
from bvae_rt.engine import GraphedTrainStep, StagedBatch, TrainEngine, cosine_lr
from models.beta_vae import BetaVAE
from training.schedulers import BetaScheduler, CapacityScheduler
from utils.brain_tumor_utils.config_parser import get_config, set_config
from utils.brain_tumor_utils.io import ensure_dirs, load_sharded_checkpoint, save_sharded_checkpoint
from utils.brain_tumor_utils.logger import init_logger, log_config, log_metrics


def set_seeds(seed):
    random.seed(seed)
    np.random.seed(seed)
    torch.manual_seed(seed)


def weight_init(m):
    """Kaiming-normal weights, zero biases on every conv / linear layer (reference: train.py:217-221)."""
    if isinstance(m, (torch.nn.Conv2d, torch.nn.ConvTranspose2d, torch.nn.Linear)):
        torch.nn.init.kaiming_normal_(m.weight, nonlinearity="relu")
        if m.bias is not None:
            torch.nn.init.zeros_(m.bias)


class SyntheticLoader:
    """``n_batches`` device-resident batches of the demo-data distribution (bvae_rt.data), dict-shaped like the
    reference dataset's samples (``image``, ``label``)."""

    def __init__(self, n_batches, batch, size, in_ch, device, seed):
        self.batches = [synthetic_batch(batch, size, seed=seed + i, device=device, in_ch=in_ch) for i in range(n_batches)]
        self.labels = [[j % 4 for j in range(batch)] for _ in range(n_batches)]

    def __iter__(self):
        for x, y in zip(self.batches, self.labels):
            yield {"image": x, "label": y}

    def __len__(self):
        return len(self.batches)


class EarlyStopping:
    def __init__(self, patience=20, min_delta=0.0):
        self.patience, self.min_delta = patience, min_delta
        self.best, self.bad, self.should_stop = None, 0, False

    def update(self, value):
        if self.best is None or value < self.best - self.min_delta:
            self.best, self.bad = value, 0
        else:
            self.bad += 1
            self.should_stop = self.bad >= self.patience


def _lr_at(cfg, epochs, epoch_index, total_steps):
    """Learning rate the reference's scheduler would have set: cosine steps once per EPOCH with T_max = epochs
    (train.py:375, callbacks.py:108-111); 'step' is StepLR(30, 0.5) stepped once per BATCH (train.py:332-333)."""
    base = cfg.optimization.lr
    kind = str(cfg.optimization.scheduler).lower()
    if kind == "none":
        return base
    if kind == "cosine":
        return cosine_lr(base, epoch_index, epochs)
    if kind == "step":
        return base * (0.5 ** (total_steps // 30))
    raise ValueError("unsupported scheduler")


def _payload(model, engine, epoch, total_steps, val_total):
    return {
        "epoch": epoch,
        "total_steps": total_steps,
        "model_state": {k: v.detach().cpu().clone() for k, v in model.state_dict().items()},
        "optim_state": engine.optimizer_state_dict(),
        "val_total": val_total,
    }


def _staged_images(loader, step_fn, augment=None):
    """The loader's image batches, one step ahead: when batch i is handed to the step, the host->device copy of batch
    i+1 is already running on the copy stream (GraphedTrainStep.prefetch).  Without a graphed step, or for batches
    that already live on the device, the tensors pass through unchanged.  A batch whose shape differs from the graph's
    static input (the last, partial batch of an epoch: the reference's DataLoader has no ``drop_last``,
    datautils.py:78-91) is not staged either; ``train`` runs it through the un-captured step.
    ``augment``: uint8 batches are the dataset's bytes BEFORE the train transform; they get the reference's flip /
    rotation / brightness on the device (bvae_rt.augment).  Float batches come from a loader that already ran its own
    transform (ToTensor is its last stage) and pass through untouched."""
    prev = None
    for batch_d in loader:
        x = batch_d["image"]
        if isinstance(x, (list, tuple)):
            x = torch.stack(x, 0)
        if step_fn is not None and not x.is_cuda and tuple(x.shape) == tuple(step_fn.x.shape):
            x = step_fn.prefetch(x, augment)
        elif augment is not None and x.dtype == torch.uint8:
            x = augment(x.to(step_fn.x.device if step_fn is not None else torch.cuda.current_device(), non_blocking=True))
        if prev is not None:
            yield prev
        prev = x
    if prev is not None:
        yield prev


def train(resume="none", cfg=None, train_loader=None, test_loader=None, use_graph=True):
    if cfg is not None:
        set_config(cfg)                  # the helpers below read the process-wide singleton, as the reference's do
    cfg = get_config()
    distributed = dist.is_available() and dist.is_initialized()
    rank = dist.get_rank() if distributed else 0
    if not torch.cuda.is_available():
        raise RuntimeError("training.train (bvae_b200) needs a CUDA device; there is no CPU path")
    device = torch.device("cuda", int(os.environ.get("LOCAL_RANK", "0")) if distributed else torch.cuda.current_device())
    torch.cuda.set_device(device)
    if rank == 0:
        ensure_dirs()
        init_logger()
        log_config()
    set_seeds(cfg.data.seed)

    debug_cfg = getattr(cfg, "debug", None)
    debug_enabled = bool(debug_cfg and debug_cfg.enabled)
    epochs = debug_cfg.epochs if debug_enabled else cfg.training.epochs
    batch = cfg.training.batch_size
    size = cfg.data.image_size
    in_ch = 1 if cfg.data.grayscale else 3
    if train_loader is None:
        n_train = max(1, (debug_cfg.train_samples if debug_enabled else 32 * batch) // batch)
        train_loader = SyntheticLoader(n_train, batch, size, in_ch, device, seed=1000 * rank)
    if test_loader is None:
        n_test = max(1, (debug_cfg.test_samples if debug_enabled else 8 * batch) // batch)
        test_loader = SyntheticLoader(n_test, batch, size, in_ch, device, seed=500000 + 1000 * rank)

    model = BetaVAE(cfg)
    model.apply(weight_init)            # same RNG consumption order as the reference -> identical initial weights
    model.to(device)
    opt = str(cfg.optimization.optimizer).lower()
    if opt not in ("adam", "adamw"):
        raise NotImplementedError(f"optimizer '{opt}': the fused step implements adam / adamw")
    clip = cfg.training.grad_clip if cfg.training.grad_clip > 0 else 0.0
    engine = TrainEngine(model, lr=cfg.optimization.lr, weight_decay=cfg.optimization.weight_decay,
                         decoupled_wd=(opt == "adamw"), clip=clip)
    mixed = bool(cfg.training.mixed_precision)
    step_fn = GraphedTrainStep(engine, (batch, in_ch, size, size), autocast_dtype=torch.bfloat16 if mixed else None,
                               seed=cfg.data.seed + rank) if use_graph else None
    beta_scheduler = BetaScheduler(cfg, total_epochs=epochs)
    capacity_scheduler = CapacityScheduler(cfg, total_epochs=epochs)
    loss_cfg = getattr(cfg, "loss", None)
    capacity_weight_cfg = getattr(loss_cfg, "capacity_weight", None) if loss_cfg else None
    # the reference builds its train transform from cfg.augmentation (augmentations.py:5-18); test data is never augmented
    augment = DeviceAugment.from_config(cfg) if hasattr(cfg, "augmentation") else None
    early = EarlyStopping(patience=20)
    models_dir, run_id = cfg.paths.models_dir, cfg.paths.run_id
    best_value = None

    start_epoch, total_steps = 1, 0
    if resume in ("best", "latest"):
        path = os.path.join(models_dir, f"{run_id}_{resume}.pt")
        try:
            payload = load_sharded_checkpoint(path, map_location=device)
        except FileNotFoundError:
            payload = None
            print(f"[RESUME] Requested '{resume}' but checkpoint not found at {path}; starting fresh.")
        if payload is not None:
            model.load_state_dict(payload.get("model_state", payload))
            model.invalidate_packs()
            if "optim_state" in payload:
                engine.load_optimizer_state_dict(payload["optim_state"])
            start_epoch = payload.get("epoch", 0) + 1
            total_steps = payload.get("total_steps", 0)
            print(f"[RESUME] Loaded checkpoint '{resume}' from {path}, restarting at epoch {start_epoch}")

    history = []
    for epoch in range(start_epoch, epochs + 1):
        beta = beta_scheduler.value(epoch - 1)
        capacity = capacity_scheduler.value(epoch) if capacity_scheduler.enabled else None
        free_bits = 0.0
        if capacity is None and loss_cfg is not None and hasattr(loss_cfg, "free_bits"):
            free_bits = loss_cfg.free_bits
        kw = dict(beta=float(beta), capacity=capacity, free_bits=float(free_bits),
                  capacity_weight=capacity_weight_cfg if capacity is not None else None)
        model.set_beta(beta)
        model.train()
        running = torch.zeros(6, device=device)     # packed order: total, recon, kl, recon_ffl, recon_base, recon_lpips
        last = None
        n_batches = 0
        for i, x in enumerate(_staged_images(train_loader, step_fn, augment)):
            engine.lr = _lr_at(cfg, epochs, epoch - 1, total_steps)
            full = isinstance(x, StagedBatch) or tuple(x.shape) == (batch, in_ch, size, size)
            if step_fn is not None and full:
                losses = step_fn(x, **kw)
            else:
                # no graph, or a ragged (partial) batch: the same step enqueued kernel by kernel at the batch's own size
                x = x.to(device, non_blocking=True)
                engine.zero_grad()
                with torch.autocast("cuda", dtype=torch.bfloat16, enabled=mixed):
                    losses = model.loss(x, **kw)
                losses["total"].backward()
                engine.step()
                # plain tensors from here on: a live autograd graph would keep this step's AccumulateGrad nodes (and
                # their stream) alive and invalidate the next graph capture (GraphedTrainStep._body does the same)
                losses = {k: (v.detach() if isinstance(v, torch.Tensor) else v) for k, v in losses.items()}
            running += losses["packed"]             # the model's six logged scalars as one device vector
            last = losses
            n_batches = i + 1
            total_steps += 1
            if rank == 0 and total_steps % cfg.logging.log_every_n_steps == 0:
                r = (running / n_batches).tolist()          # the one device->host read of this logging interval
                log_metrics({
                    "epoch": epoch, "beta": float(beta), "capacity": float(capacity) if capacity is not None else 0.0,
                    "train_total_loss": r[0], "train_recon_loss": r[1], "train_recon_base": r[4],
                    "train_recon_lpips": r[5], "train_recon_ffl": r[3], "train_kl": r[2], "train_kl_mean": r[2],
                    "train_kl_effective_last": float(last["kl_effective"]),
                    "train_kl_per_dim_mean": float(last["kl_per_dim"].mean()), "loss_mode": last["mode"],
                    "mu_mean_batch": float(last["mu"].mean()), "z_std_batch": float(last["z"].std()),
                    "lr": engine.lr,
                }, step=total_steps, phase="train")
            if debug_enabled and i + 1 >= debug_cfg.max_train_batches:
                break
        train_means = (running / max(1, n_batches)).tolist()

        # ---- validation: fp32, no autocast (reference train.py:386-413)
        model.eval()
        vsum = torch.zeros(6, device=device)
        vb = 0
        vlast = None
        with torch.no_grad():
            for j, batch_d in enumerate(test_loader):
                x = batch_d["image"]
                if isinstance(x, (list, tuple)):
                    x = torch.stack(x, 0)
                l = model.loss(x.to(device, non_blocking=True), **kw)
                vsum += l["packed"]
                vlast = l
                vb += 1
                if debug_enabled and j + 1 >= debug_cfg.max_val_batches:
                    break
        if distributed:
            # every rank validates its own shard; the monitored value (checkpoint selection, early stopping) must be the
            # same number on all ranks or they leave the epoch loop at different times and the next captured
            # all-reduce never completes
            cnt = torch.tensor([float(vb)], device=device)
            dist.all_reduce(vsum)
            dist.all_reduce(cnt)
            v = (vsum / cnt.clamp_min(1.0)).tolist()
        else:
            v = (vsum / max(1, vb)).tolist()
        val_total = v[0]
        nan = float("nan")
        if rank == 0:
            log_metrics({
                "epoch": epoch, "beta": float(beta), "capacity": float(capacity) if capacity is not None else 0.0,
                "val_total_loss": v[0], "val_recon_loss": v[1], "val_recon_base": v[4], "val_recon_lpips": v[5],
                "val_recon_ffl": v[3], "val_kl": v[2],
                "val_kl_per_dim_mean": float(vlast["kl_per_dim"].mean()) if vlast is not None else nan,
                "loss_mode": vlast["mode"] if vlast is not None else None,
                "train_kl_mean": train_means[2],
                "train_kl_effective_last": float(last["kl_effective"]) if last is not None else nan,
                # the sklearn linear probe of the reference (train.py:24-96) is analysis code, out of scope
                "latent_probe_auc": nan, "best_dim_auc": nan, "best_dim_corr": nan, "best_dim_r2": nan,
            }, step=total_steps, phase="val")
            latest = os.path.join(models_dir, f"{run_id}_latest.pt")
            save_sharded_checkpoint(latest, _payload(model, engine, epoch, total_steps, val_total), num_shards=2)
            if best_value is None or val_total < best_value:
                best_value = val_total
                save_sharded_checkpoint(os.path.join(models_dir, f"{run_id}_best.pt"),
                                        _payload(model, engine, epoch, total_steps, val_total), num_shards=2)
        history.append({"epoch": epoch, "train_total": train_means[0], "val_total": val_total, "lr": engine.lr,
                        "beta": float(beta), "capacity": capacity, "steps": total_steps})
        early.update(val_total if math.isfinite(val_total) else float("inf"))
        if early.should_stop:
            break
    engine.close()
    return history


def _parse_args():
    ap = argparse.ArgumentParser()
    ap.add_argument("--config", type=str, default=None, help="YAML config (sets CONFIG_PATH)")
    ap.add_argument("--resume", type=str, default="none", choices=["none", "best", "latest"])
    return ap.parse_args()


if __name__ == "__main__":
    args = _parse_args()
    if args.config:
        os.environ["CONFIG_PATH"] = args.config
    if int(os.environ.get("WORLD_SIZE", "1")) > 1:
        torch.cuda.set_device(int(os.environ.get("LOCAL_RANK", "0")))
        dist.init_process_group("nccl")
    train(resume=args.resume)
