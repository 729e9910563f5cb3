"""Sharded checkpoint files in the reference's layout (reference: src/utils/brain_tumor_utils/io.py:8-70).

Contract: for a base path ``<root>.pt`` the files are ``<root>_shard{i}.pt``; the sorted
``model_state`` keys are dealt round-robin (key number ``idx`` goes to shard ``idx % num_shards``); every shard is a
``torch.save``d dict holding a full copy of all other payload fields plus ``shard_id`` / ``num_shards``; a stale
un-sharded base file is removed.  Loading globs the shards, merges ``model_state`` and takes the remaining fields from
the first shard; a plain single-file checkpoint is accepted as a fallback.  Files written here load with the
reference's ``load_sharded_checkpoint`` and vice versa."""
import glob
import json
import os

import torch

from .config_parser import get_config


def _split_ext(base_path):
    root, ext = os.path.splitext(base_path)
    return root, (ext or ".pt")


def shard_paths(base_path, num_shards):
    root, ext = _split_ext(base_path)
    return [f"{root}_shard{i}{ext}" for i in range(num_shards)]


def save_sharded_checkpoint(base_path, payload, num_shards=2):
    if payload.get("model_state") is None:
        raise ValueError("payload missing model_state for sharded checkpoint save")
    os.makedirs(os.path.dirname(base_path) or ".", exist_ok=True)
    num_shards = max(1, int(num_shards))
    state = payload["model_state"]
    paths = shard_paths(base_path, num_shards)
    ordered = sorted(state.keys())
    for shard_id, path in enumerate(paths):
        part = dict(payload)
        part["model_state"] = {k: state[k] for k in ordered[shard_id::num_shards]}
        part["shard_id"] = shard_id
        part["num_shards"] = num_shards
        torch.save(part, path)
    if os.path.exists(base_path):
        os.remove(base_path)
    return paths


def load_sharded_checkpoint(base_path, map_location=None, num_shards=None):
    root, ext = _split_ext(base_path)
    paths = sorted(glob.glob(f"{root}_shard*{ext}")) if num_shards is None else shard_paths(base_path, num_shards)
    if paths and all(os.path.exists(p) for p in paths):
        merged, meta = {}, None
        for p in paths:
            part = torch.load(p, map_location=map_location, weights_only=False)
            merged.update(part.get("model_state", {}))
            if meta is None:
                meta = {k: v for k, v in part.items() if k != "model_state"}
        meta["model_state"] = merged
        return meta
    if os.path.exists(base_path):
        return torch.load(base_path, map_location=map_location, weights_only=False)
    raise FileNotFoundError(f"No checkpoint found at {base_path} or shards")


def ensure_dirs():
    """Create the output directories of the active config (reference: io.py:72-78, same no-argument call)."""
    cfg = get_config()
    for key in ("outputs_dir", "models_dir", "figures_dir", "tables_dir"):
        os.makedirs(getattr(cfg.paths, key), exist_ok=True)
    os.makedirs(os.path.join(cfg.paths.outputs_dir, "logs"), exist_ok=True)


def run_artifact_dir():
    cfg = get_config()
    d = os.path.join(cfg.paths.outputs_dir, cfg.paths.run_id)
    os.makedirs(d, exist_ok=True)
    return d


def model_checkpoint_path(epoch=None, tag=None):
    """``<models_dir>/<run_id>_{tag | epoch<N> | latest}.pt`` (reference: io.py:86-94)."""
    cfg = get_config()
    os.makedirs(cfg.paths.models_dir, exist_ok=True)
    stem = tag if tag else (f"epoch{epoch}" if epoch is not None else "latest")
    return os.path.join(cfg.paths.models_dir, f"{cfg.paths.run_id}_{stem}.pt")


def save_checkpoint(model, optimizer, epoch, metrics=None, tag=None, extra=None):
    path = model_checkpoint_path(None if tag else epoch, tag=tag)
    save_sharded_checkpoint(path, {
        "epoch": epoch,
        "model_state": model.state_dict(),
        "optimizer_state": None if optimizer is None else optimizer.state_dict(),
        "metrics": metrics or {},
        "extra": extra or {},
    }, num_shards=2)
    return path


def save_json(data, name):
    out = os.path.join(get_config().paths.outputs_dir, f"{name}.json")
    with open(out, "w") as f:
        json.dump(data, f, indent=2)
    return out


def save_table(df, name):
    path = os.path.join(get_config().paths.tables_dir, f"{name}.csv")
    df.to_csv(path, index=False)
    return path


def save_figure(fig, name):
    path = os.path.join(get_config().paths.figures_dir, f"{name}.png")
    fig.savefig(path, bbox_inches="tight")
    return path
