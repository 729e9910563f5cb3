"""``METRICS {json}`` / ``CONFIG {json}`` log lines, the format the reference's plotting scripts re-parse
(reference: src/utils/brain_tumor_utils/logger.py:10-43)."""
import json
import logging
import os
import sys

_logger = None


def init_logger(cfg=None):
    global _logger
    if _logger is not None:
        return _logger
    logger = logging.getLogger("bvae_b200")
    logger.setLevel(logging.INFO)
    logger.propagate = False
    fmt = logging.Formatter("%(asctime)s %(levelname)s %(message)s")
    sh = logging.StreamHandler(sys.stdout)
    sh.setFormatter(fmt)
    logger.addHandler(sh)
    if cfg is not None and getattr(cfg.logging, "log_to_file", False):
        log_dir = os.path.join(cfg.paths.outputs_dir, "logs")
        os.makedirs(log_dir, exist_ok=True)
        fh = logging.FileHandler(os.path.join(log_dir, f"{cfg.paths.run_id}.log"))
        fh.setFormatter(fmt)
        logger.addHandler(fh)
    _logger = logger
    return logger


def log_config(cfg):
    init_logger(cfg).info("CONFIG " + json.dumps(cfg.to_dict()))


def log_metrics(metrics, step=None, phase="train"):
    payload = {"phase": phase, "step": step}
    payload.update(metrics)
    init_logger().info("METRICS " + json.dumps(payload))
