"""``METRICS {json}`` / ``CONFIG {json}`` log lines, the format the reference's plotting scripts re-parse.

Same call signatures as the reference -- ``init_logger(name="beta_vae_se")``, ``log_config()``,
``log_metrics(metrics, step=None, phase="train")`` (reference: src/utils/brain_tumor_utils/logger.py:10,34,39) -- so the
reference's ``train.py:226-227`` works on top of this module; the config always comes from the ``get_config()``
singleton, as there."""
import json
import logging
import os
import sys

from .config_parser import get_config

_logger = None


def init_logger(name="beta_vae_se"):
    global _logger
    if _logger is not None:
        return _logger
    cfg = get_config()
    logger = logging.getLogger(name)
    logger.setLevel(getattr(logging, str(cfg.logging.log_level).upper(), logging.INFO))
    logger.propagate = False
    if not logger.handlers:
        fmt = logging.Formatter("%(asctime)s | %(levelname)s | %(message)s")
        out = logging.StreamHandler(sys.stdout)
        out.setFormatter(fmt)
        logger.addHandler(out)
        if cfg.logging.log_to_file:
            log_dir = os.path.join(cfg.paths.outputs_dir, "logs")
            os.makedirs(log_dir, exist_ok=True)
            to_file = logging.FileHandler(os.path.join(log_dir, f"{cfg.paths.run_id}.log"))
            to_file.setFormatter(fmt)
            logger.addHandler(to_file)
    _logger = logger
    return logger


def log_config():
    init_logger().info("CONFIG " + json.dumps(get_config().to_dict()))


def log_metrics(metrics, step=None, phase="train"):
    record = {"phase": phase, "step": step}
    record.update(metrics)
    init_logger().info("METRICS " + json.dumps(record))
