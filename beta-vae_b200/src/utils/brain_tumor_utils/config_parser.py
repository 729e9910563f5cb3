"""YAML config -> frozen attribute tree, process-wide singleton.  Same keys, same resolution
order (argument, then $CONFIG_PATH, then the two default file names) and same validation as
the reference (reference: src/utils/brain_tumor_utils/config_parser.py:6-90), so existing
YAML files and ``get_config()`` call sites work unchanged."""
import os
import threading
from copy import deepcopy

import yaml

REQUIRED_SECTIONS = (
    "paths", "data", "model", "training", "optimization", "beta_schedule",
    "augmentation", "evaluation", "inference", "logging", "experiment", "debug",
)
DEFAULT_CANDIDATES = ("configs/beta_vae_se.yaml", "configs/overfit_capacity.yaml")


class _Frozen:
    """Read-only attribute view of a (nested) dict; re-assignment raises."""

    def __init__(self, mapping):
        for key, value in mapping.items():
            object.__setattr__(self, key, _Frozen(value) if isinstance(value, dict) else value)

    def __setattr__(self, key, value):
        if key in self.__dict__:
            raise AttributeError("Frozen config is immutable")
        object.__setattr__(self, key, value)

    def __getitem__(self, item):
        return getattr(self, item)

    def to_dict(self):
        return {k: (v.to_dict() if isinstance(v, _Frozen) else v) for k, v in self.__dict__.items()}


def freeze(mapping):
    return _Frozen(deepcopy(mapping))


def _validate(raw):
    missing = [k for k in REQUIRED_SECTIONS if k not in raw]
    if missing:
        raise ValueError(f"Missing required top-level keys: {missing}")
    if raw["data"]["class_mode"] not in ("binary", "multiclass"):
        raise ValueError("data.class_mode must be binary or multiclass")
    if raw["beta_schedule"]["type"] not in ("constant", "linear", "cyclical"):
        raise ValueError("beta_schedule.type invalid")
    return True


def _resolve_config_path(path=None):
    tried = []
    for cand in (path, os.environ.get("CONFIG_PATH"), *DEFAULT_CANDIDATES):
        if not cand:
            continue
        cand = os.path.expanduser(str(cand))
        tried.append(cand)
        if os.path.exists(cand):
            return cand
    raise FileNotFoundError(f"Config file not found. Set CONFIG_PATH or pass a path. Tried: {tried}")


def load_config(path=None):
    with open(_resolve_config_path(path), "r") as f:
        raw = yaml.safe_load(f)
    _validate(raw)
    return raw


_config_cache = None
_config_lock = threading.Lock()


def get_config(path=None):
    global _config_cache
    if _config_cache is None:
        with _config_lock:
            if _config_cache is None:
                _config_cache = freeze(load_config(path))
    return _config_cache


def set_config(cfg):
    """Install ``cfg`` (a mapping or an already frozen tree) as the process-wide singleton -- for callers that build
    a config in code (tests, sweeps) instead of pointing ``CONFIG_PATH`` at a file.  Not in the reference."""
    global _config_cache
    with _config_lock:
        _config_cache = cfg if isinstance(cfg, _Frozen) else freeze(cfg)
    return _config_cache
