"""The training loop around the engine (GPU): debug config on synthetic batches -- the reference's own plumbing check
 -- checkpoint files in the reference layout, resume, and the bf16 CUDA-graph path."""
import copy
import glob
import os

import pytest
import torch

import cfgs

pytestmark = pytest.mark.gpu


def _cfg(tmp_path, **over):
    from utils.brain_tumor_utils.config_parser import freeze
    raw = copy.deepcopy(cfgs.shipped("beta_vae_se_debug"))
    for k in ("figures_dir", "models_dir", "outputs_dir", "tables_dir"):
        raw["paths"][k] = str(tmp_path / raw["paths"][k])
    raw["model"]["base_channels"] = 16           # keep the test quick; layer plan otherwise as shipped
    for k, v in over.items():
        sec, key = k.split("__")
        raw[sec][key] = v
    return freeze(raw)


def test_debug_config_trains_checkpoints_and_resumes(tmp_path):
    from training.train import train
    from utils.brain_tumor_utils.io import load_sharded_checkpoint
    from models.beta_vae import BetaVAE
    cfg = _cfg(tmp_path)
    hist = train(cfg=cfg)                                         # fp32 (mixed_precision: false), l1 + FFL + free bits
    assert [h["epoch"] for h in hist] == [1, 2] and hist[-1]["steps"] == 6          # 3 train batches per epoch
    assert all(torch.isfinite(torch.tensor([h["train_total"], h["val_total"]])).all() for h in hist)
    assert hist[0]["beta"] == pytest.approx(0.05) and hist[1]["beta"] == pytest.approx(0.05 + 0.1 / 25)
    assert hist[0]["lr"] == pytest.approx(2e-3) and hist[1]["lr"] == pytest.approx(1e-3)   # cosine, T_max = 2 epochs
    files = sorted(os.path.basename(p) for p in glob.glob(os.path.join(cfg.paths.models_dir, "*.pt")))
    assert files == ["beta_vae_se_debug_best_shard0.pt", "beta_vae_se_debug_best_shard1.pt",
                     "beta_vae_se_debug_latest_shard0.pt", "beta_vae_se_debug_latest_shard1.pt"]
    ck = load_sharded_checkpoint(os.path.join(cfg.paths.models_dir, "beta_vae_se_debug_latest.pt"), map_location="cpu")
    assert ck["epoch"] == 2 and ck["total_steps"] == 6 and "optim_state" in ck and "val_total" in ck
    fresh = BetaVAE(cfg)
    fresh.load_state_dict(ck["model_state"], strict=True)         # same 72 keys / shapes / fp32 as the reference
    assert all(v.dtype == torch.float32 for v in ck["model_state"].values())
    st = ck["optim_state"]
    assert len(st["state"]) == len(list(fresh.parameters())) and int(st["state"][0]["step"]) == 6
    # torch.optim.Adam accepts the engine's optimizer state (what the reference's resume does, train.py:270-272)
    torch.optim.Adam(fresh.parameters(), lr=1e-3).load_state_dict(st)
    assert train(resume="latest", cfg=cfg) == []                  # nothing left to do after epoch 2
    longer = _cfg(tmp_path, debug__epochs=3)
    more = train(resume="latest", cfg=longer)
    assert [h["epoch"] for h in more] == [3] and more[0]["steps"] == 9


def test_mixed_precision_graph_path_learns(tmp_path):
    from training.train import train
    cfg = _cfg(tmp_path, training__mixed_precision=True, debug__epochs=4, debug__max_train_batches=3,
               model__reconstruction_loss="mse", loss__free_bits=0.0)
    hist = train(cfg=cfg)                                         # 2 eager steps, then CUDA-graph replays (bf16 kernels)
    assert len(hist) == 4
    assert hist[-1]["train_total"] < hist[0]["train_total"]
