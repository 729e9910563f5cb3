"""Sharded checkpoint files in the reference's layout (reference: src/utils/brain_tumor_utils/io.py:8-70).

Contract: for a base path ``<root>.pt`` the files are ``<root>_shard{i}.pt``; the sorted
``model_state`` keys are dealt round-robin (key number ``idx`` goes to shard ``idx % num_shards``); every shard is a
``torch.save``d dict holding a full copy of all other payload fields plus ``shard_id`` / ``num_shards``; a stale
un-sharded base file is removed.  Loading globs the shards, merges ``model_state`` and takes the remaining fields from
the first shard; a plain single-file checkpoint is accepted as a fallback.  Files written here load with the
reference's ``load_sharded_checkpoint`` and vice versa."""
import glob
import os

import torch


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


def ensure_dirs(cfg):
    for key in ("outputs_dir", "models_dir", "figures_dir", "tables_dir"):
        os.makedirs(getattr(cfg.paths, key), exist_ok=True)
    os.makedirs(os.path.join(cfg.paths.outputs_dir, "logs"), exist_ok=True)
