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


class _ListLoader:
    def __init__(self, batches):
        self.batches = batches

    def __iter__(self):
        return iter(self.batches)

    def __len__(self):
        return len(self.batches)


def test_ragged_tail_batch_runs_through_the_uncaptured_step(tmp_path):
    """The reference's DataLoader has no drop_last (datautils.py:78-91): an epoch of N % batch != 0 samples ends with a
    partial batch.  The captured graph has a fixed batch; the tail (here 5 and then 1 sample, host tensors) must be
    trained on at its own size, not rejected and not broadcast into the static buffer."""
    from training.train import train
    cfg = _cfg(tmp_path, training__mixed_precision=True, debug__epochs=2, debug__max_train_batches=10)
    g = torch.Generator().manual_seed(0)
    mk = lambda n: {"image": torch.rand(n, 1, 128, 128, generator=g), "label": [0] * n}     # noqa: E731
    loader = _ListLoader([mk(8), mk(8), mk(8), mk(5), mk(1)])
    hist = train(cfg=cfg, train_loader=loader, test_loader=_ListLoader([mk(8), mk(3)]))
    assert [h["steps"] for h in hist] == [5, 10]
    assert all(torch.isfinite(torch.tensor([h["train_total"], h["val_total"]])).all() for h in hist)


@pytest.mark.parametrize("prec", ["fp32", "bf16"])
def test_loss_on_uint8_images_equals_loss_on_their_floats(prec):
    """loss(u8) must use x / 255 for the encoder input AND for both targets (reconstruction, focal-frequency)."""
    from models.beta_vae import BetaVAE
    from utils.brain_tumor_utils.config_parser import freeze
    cfg = cfgs.small(loss__use_ffl=True, loss__ffl_weight=0.5)
    torch.manual_seed(1)
    m = BetaVAE(freeze(cfg), precision=prec).apply(cfgs.weight_init).cuda().train()
    u8 = torch.randint(0, 256, (6, 1, 32, 32), dtype=torch.uint8, device="cuda")
    eps = torch.randn(6, 8, device="cuda")
    kw = dict(beta=2.0, capacity=None, free_bits=0.0, capacity_weight=None)
    a = m.loss(u8, eps=eps, **kw)
    b = m.loss(u8.float() / torch.full((), 255.0, device="cuda"), eps=eps, **kw)
    for k in ("total", "recon", "recon_base", "recon_ffl", "kl_mean"):
        assert float(a[k]) == float(b[k]), (k, float(a[k]), float(b[k]))
    assert torch.equal(a["recon_img"], b["recon_img"])
    with torch.no_grad():
        m.eval()
        assert float(m.loss(u8, eps=eps, **kw)["total"]) == float(a["total"])
