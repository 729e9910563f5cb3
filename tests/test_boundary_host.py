"""Host-side checks of the drop-in boundary and of the measurement plumbing (CPU, no kernels launched):

  * the drop-in's ``models`` / ``training`` / ``utils`` packages coexist with the reference's on one module path
    (what INTEGRATION.md section 3 relies on; tests/test_gpu_integration.py then runs the reference's train.py);
  * helper signatures the reference's scripts call with no arguments (``init_logger()``, ``log_config()``,
    ``ensure_dirs()``: src/training/train.py:225-227);
  * the staged reference (oracle/_ref) is byte-identical to /root/reference, and the oracle port agrees with the LIVE
    reference class on a fresh seed (the goldens pin fixed seeds);
  * bench.py prints the same ``config`` object in both arms and names every BASELINE.json workload."""
import json
import os
import subprocess
import sys

import pytest
import torch

import cfgs

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
PKG = os.path.join(ROOT, "beta-vae_b200")
REF = os.path.join(ROOT, "oracle", "_ref")
needs_ref = pytest.mark.skipif(not os.path.exists(os.path.join(REF, "src", "models", "beta_vae.py")),
                               reason="oracle/_ref not staged (python oracle/make_ref.py)")


def _py(code, extra_path, env_over=None, cwd=None):
    env = dict(os.environ)
    env["PYTHONPATH"] = os.pathsep.join(extra_path)
    env.update(env_over or {})
    p = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True, env=env, cwd=cwd, timeout=600)
    assert p.returncode == 0, p.stdout[-2000:] + p.stderr[-2000:]
    return p.stdout


@needs_ref
def test_drop_in_packages_span_the_reference_tree(tmp_path):
    code = (
        "import training.callbacks as a, training.schedulers as b, training.losses as c, models.beta_vae as d\n"
        "import utils.brain_tumor_utils.datautils as e, utils.brain_tumor_utils.io as f, utils.brain_tumor_utils.logger as g\n"
        "import data_processing.augmentations as h, utils.brain_tumor_utils.config_parser as i\n"
        "from utils.brain_tumor_utils import get_config, build_dataloaders, init_logger, save_checkpoint\n"
        "import json; print(json.dumps({m.__name__: m.__file__ for m in (a, b, c, d, e, f, g, h, i)}))\n")
    out = _py(code, [os.path.join(PKG, "src"), PKG, os.path.join(REF, "src")],
              {"CONFIG_PATH": os.path.join(PKG, "configs", "beta_vae_se_debug.yaml")}, cwd=str(tmp_path))
    where = json.loads(out.strip().splitlines()[-1])
    ours = {"training.schedulers", "training.losses", "models.beta_vae", "utils.brain_tumor_utils.io",
            "utils.brain_tumor_utils.logger", "utils.brain_tumor_utils.config_parser"}
    for name, path in where.items():
        want = PKG if name in ours else REF
        assert os.path.realpath(path).startswith(os.path.realpath(want)), (name, path)


def test_helper_signatures_match_the_reference_call_sites(tmp_path, monkeypatch):
    import inspect
    from utils.brain_tumor_utils import config_parser, io, logger
    assert list(inspect.signature(logger.init_logger).parameters) == ["name"]
    assert inspect.signature(logger.init_logger).parameters["name"].default == "beta_vae_se"
    assert list(inspect.signature(logger.log_config).parameters) == []
    assert list(inspect.signature(logger.log_metrics).parameters) == ["metrics", "step", "phase"]
    assert list(inspect.signature(io.ensure_dirs).parameters) == []
    assert list(inspect.signature(io.save_checkpoint).parameters) == ["model", "optimizer", "epoch", "metrics", "tag", "extra"]
    raw = cfgs.shipped("beta_vae_se_debug")
    raw["paths"] = {k: (str(tmp_path / v) if k.endswith("_dir") else v) for k, v in raw["paths"].items()}
    raw["logging"]["log_to_file"] = True
    monkeypatch.setattr(config_parser, "_config_cache", None)
    monkeypatch.setattr(logger, "_logger", None)
    config_parser.set_config(raw)
    io.ensure_dirs()
    logger.log_config()
    logger.log_metrics({"train_total_loss": 1.5}, step=3, phase="train")
    log = (tmp_path / "outputs" / "logs" / f"{raw['paths']['run_id']}.log").read_text()
    assert "CONFIG {" in log and 'METRICS {"phase": "train", "step": 3, "train_total_loss": 1.5}' in log
    assert " | INFO | " in log                      # the reference's line format (logger.py:21)
    assert io.model_checkpoint_path(tag="best").endswith(f"{raw['paths']['run_id']}_best.pt")
    assert io.model_checkpoint_path(epoch=4).endswith("_epoch4.pt")
    for h in list(logger._logger.handlers):
        h.close()


@needs_ref
def test_staged_reference_is_byte_identical():
    sys.path.insert(0, ROOT)
    from oracle import make_ref
    assert make_ref.check() == 0


@needs_ref
def test_oracle_port_matches_the_live_reference_class_on_a_fresh_seed(tmp_path):
    """The goldens pin fixed seeds; this runs the staged, unmodified class next to the oracle port on another one."""
    code = f"""
import sys, json, torch
sys.path.insert(0, {ROOT!r}); sys.path.insert(0, {os.path.join(ROOT, 'tests')!r})
from oracle import beta_vae_oracle as O
import cfgs
from models.beta_vae import BetaVAE
torch.manual_seed(777)
m = BetaVAE(); m.apply(cfgs.weight_init); m.train()
g = torch.Generator().manual_seed(31)
x = torch.rand(3, 1, 128, 128, generator=g); eps = torch.randn(3, 32, generator=g)
m.reparameterize = lambda mu, logvar: mu + eps * torch.exp(0.5 * logvar)
kw = dict(beta=0.05, capacity=None, free_bits=0.5, capacity_weight=None)
out = m.loss(x, **kw); out['total'].backward()
sd = {{k: v.detach().clone() for k, v in m.state_dict().items()}}
spec = O.spec_from_config(m.cfg)
o, og = O.loss_and_grads(sd, spec, x, eps, **kw)
worst = max(float((p.grad - og[k]).norm() / (og[k].norm() + 1e-30)) for k, p in m.named_parameters())
print(json.dumps({{'ref_total': float(out['total']), 'oracle_total': float(o['total']), 'worst_grad_rel': worst,
                  'kl_ref': float(out['kl_mean']), 'kl_oracle': float(o['kl_mean'])}}))
"""
    out = _py(code, [os.path.join(REF, "src")], {"CONFIG_PATH": os.path.join(PKG, "configs", "beta_vae_se_debug.yaml")},
              cwd=str(tmp_path))
    r = json.loads(out.strip().splitlines()[-1])
    assert abs(r["ref_total"] - r["oracle_total"]) <= 1e-5 * abs(r["ref_total"]), r
    assert abs(r["kl_ref"] - r["kl_oracle"]) <= 1e-5 * abs(r["kl_ref"]), r
    assert r["worst_grad_rel"] <= 5e-4, r          # fp32 vs fp32, different summation orders on ill-conditioned tensors


def test_bench_config_is_identical_in_both_arms_and_names_every_workload():
    sys.path.insert(0, ROOT)
    import bench
    assert set(bench.WORKLOADS) == {"default", "overfit_capacity", "overfit_capacity_wide128", "inference"}
    for wl in bench.WORKLOADS:
        cfg = bench.load_cfg(wl)
        key = bench.config_key(wl, cfg, 32)
        assert set(key) == {"workload", "image_size", "per_gpu_batch", "objective"}      # no arm-specific keys
    assert bench.load_cfg("overfit_capacity_wide128")["model"]["base_channels"] == 128
    assert bench.load_cfg("overfit_capacity")["model"]["base_channels"] == 16
    # the default workload's objective is the shipped schedule's first-epoch regime (capacity mode, gamma = 1)
    from training.schedulers import BetaScheduler, CapacityScheduler
    from utils.brain_tumor_utils.config_parser import freeze
    cfg = freeze(bench.load_cfg("default"))
    assert CapacityScheduler(cfg, total_epochs=cfg.training.epochs).enabled
    assert BetaScheduler(cfg, total_epochs=cfg.training.epochs).value(0) == bench.WORKLOADS["default"]["loss_kw"]["beta"]
    assert cfg.loss.capacity_weight == bench.WORKLOADS["default"]["loss_kw"]["capacity_weight"]
    cfg = freeze(bench.load_cfg("overfit_capacity"))
    assert not CapacityScheduler(cfg, total_epochs=cfg.training.epochs).enabled
    assert BetaScheduler(cfg, total_epochs=cfg.training.epochs).value(0) == bench.WORKLOADS["overfit_capacity"]["loss_kw"]["beta"]


def test_no_product_module_touches_the_oracle_or_the_reference():
    bad = []
    for base, _, files in os.walk(PKG):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                txt = open(os.path.join(base, f), errors="replace").read()
                if "import oracle" in txt or "from oracle" in txt or "/root/reference" in txt or "oracle/_ref" in txt:
                    bad.append(os.path.join(base, f))
    assert not bad, bad


def test_stand_alone_se_and_cpu_ffl_fail_loudly():
    from models.se_blocks import SEBlock
    from training.losses import FocalFrequencyLoss
    with pytest.raises(NotImplementedError):
        SEBlock(16, 4)(torch.zeros(1, 16, 4, 4))
    with pytest.raises(RuntimeError):
        FocalFrequencyLoss()(torch.zeros(1, 1, 8, 8), torch.zeros(1, 1, 8, 8))
