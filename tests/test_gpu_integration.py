"""INTEGRATION.md section 3, executed: the reference's OWN, UNMODIFIED ``src/training/train.py`` (staged by
oracle/make_ref.py) trains on the drop-in model.

The drop-in's ``src`` goes first on the module path, the reference's ``src`` after it; ``models.beta_vae``,
``training.losses`` / ``schedulers``, ``utils.brain_tumor_utils.{config_parser,logger,io}`` resolve to this repo,
everything else (``training.callbacks``, ``utils.brain_tumor_utils.datautils``, ``data_processing.augmentations``)
to the reference -- the packages span both trees (``pkgutil.extend_path``).  The reference's demo-data generator
(scripts/generate_demo_data.py) writes the PNG dataset, then train.py runs the shipped debug config
(``use_lpips: false``: the lpips package is not available offline): 2 epochs x 3 batches of 8, torch.optim.Adam,
GradScalerWrapper, CheckpointManager, sample_reconstructions -- all the reference's own code -- with every model
forward / backward going through libbvae_b200.so.  A second run switches ``training.mixed_precision`` on, so the
reference's fp16 autocast + GradScaler drive the bf16 kernels.
Reference call sites: src/training/train.py:223-470, src/training/callbacks.py:7-90."""
import glob
import os
import subprocess
import sys

import pytest
import torch
import yaml

pytestmark = pytest.mark.gpu

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
PKG = os.path.join(ROOT, "beta-vae_b200")


def _run(cmd, cwd, env):
    p = subprocess.run(cmd, cwd=cwd, env=env, capture_output=True, text=True, timeout=900)
    assert p.returncode == 0, f"{cmd} failed:\n{p.stdout[-3000:]}\n{p.stderr[-3000:]}"
    return p.stdout + p.stderr


@pytest.mark.parametrize("mixed", [False, True], ids=["fp32_as_shipped", "mixed_precision"])
def test_reference_train_py_runs_on_the_drop_in_model(tmp_path, mixed):
    sys.path.insert(0, ROOT)
    from oracle import make_ref
    if not make_ref.available():
        pytest.skip("oracle/_ref is not staged (python oracle/make_ref.py needs /root/reference)")
    ref = make_ref.DST
    with open(os.path.join(PKG, "configs", "beta_vae_se_debug.yaml")) as f:
        cfg = yaml.safe_load(f)
    cfg["training"]["mixed_precision"] = bool(mixed)
    cfg["training"]["num_workers"] = 0
    cfg["paths"]["run_id"] = "integration_debug"
    cfg_path = tmp_path / "cfg.yaml"
    cfg_path.write_text(yaml.safe_dump(cfg))
    env = dict(os.environ)
    env["CONFIG_PATH"] = str(cfg_path)
    env["PYTHONPATH"] = os.pathsep.join([os.path.join(PKG, "src"), PKG, os.path.join(ref, "src")])
    env.pop("RANK", None)
    env.pop("WORLD_SIZE", None)

    _run([sys.executable, os.path.join(ref, "scripts", "generate_demo_data.py")], str(tmp_path), env)
    train_py = os.path.join(ref, "src", "training", "train.py")
    # run the unmodified file as __main__, then report which modules it ended up using
    code = (f"import runpy, sys\nsys.argv = ['train.py']\nrunpy.run_path({train_py!r}, run_name='__main__')\n"
            "import models.beta_vae as m, training.callbacks as c, utils.brain_tumor_utils.datautils as d\n"
            "from bvae_rt import lib\n"
            "print('USED model=%s callbacks=%s datautils=%s launches=%d' % (m.__file__, c.__file__, d.__file__, "
            "lib.query('bvae_launch_count')))\n")
    out = _run([sys.executable, "-c", code], str(tmp_path), env)

    used = [ln for ln in out.splitlines() if ln.startswith("USED ")]
    assert used, out[-2000:]
    assert os.path.join("beta-vae_b200", "src", "models", "beta_vae.py") in used[0], used[0]
    assert os.path.join("_ref", "src", "training", "callbacks.py") in used[0], used[0]
    assert os.path.join("_ref", "src", "utils", "brain_tumor_utils", "datautils.py") in used[0], used[0]
    assert int(used[0].rsplit("launches=", 1)[1]) > 500, "the run did not go through libbvae_b200.so"
    # the reference's logger lines: 2 validation records with finite losses
    import json
    vals = [json.loads(ln.split("METRICS ", 1)[1]) for ln in out.splitlines() if "METRICS " in ln]
    val = [v for v in vals if v["phase"] == "val"]
    assert len(val) == cfg["debug"]["epochs"], [v["phase"] for v in vals]
    for v in val:
        assert v["val_total_loss"] == v["val_total_loss"] and v["val_total_loss"] > 0, v
    # the reference's CheckpointManager wrote latest + best, 2 shards each, in the reference layout
    mdir = tmp_path / cfg["paths"]["models_dir"]
    for tag in ("latest", "best"):
        shards = sorted(glob.glob(str(mdir / f"integration_debug_{tag}_shard*.pt")))
        assert len(shards) == 2, (tag, os.listdir(mdir))
    sys.path.insert(0, os.path.join(PKG, "src"))
    from utils.brain_tumor_utils.io import load_sharded_checkpoint
    payload = load_sharded_checkpoint(str(mdir / "integration_debug_latest.pt"), map_location="cpu")
    assert len(payload["model_state"]) == 72 and payload["epoch"] == cfg["debug"]["epochs"]
    assert all(torch.isfinite(v).all() for v in payload["model_state"].values())
    assert os.path.exists(tmp_path / cfg["paths"]["figures_dir"] / "recon_epoch2.png")
