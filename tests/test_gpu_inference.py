"""Inference path: the drop-ins for src/inference/encode.py and generate.py against the oracle.

encode_loader must return what the reference's per-batch loop returns (same order, labels, paths; fp32 within 1e-5);
the batched edit / interpolation decodes must equal the reference's one-latent-per-iteration loop."""
import csv
import os
from types import SimpleNamespace

import numpy as np
import pytest
import torch

import cfgs
from parity import rel_err
from test_gpu_model import build_model

pytestmark = pytest.mark.gpu


def _setup(seed=3, **over):
    from oracle import beta_vae_oracle as O
    cfg = cfgs.small(**over)
    spec = O.spec_from_config(cfgs.ns(cfg))
    torch.manual_seed(seed)
    m = build_model(cfg, None, "fp32").eval()
    sd = {k: v.detach().cpu() for k, v in m.state_dict().items()}
    return O, spec, m, sd


def _loader(n_batches, bs, last, dtype):
    g = torch.Generator().manual_seed(11)
    out = []
    k = 0
    for i in range(n_batches):
        n = last if i == n_batches - 1 else bs
        u8 = torch.randint(0, 256, (n, 1, 32, 32), generator=g, dtype=torch.uint8)
        img = u8 if dtype == torch.uint8 else u8.float().div(255)
        out.append({"image": img, "label": [f"c{(k + j) % 4}" for j in range(n)], "path": [f"img_{k + j}.png" for j in range(n)]})
        k += n
    return out


@pytest.mark.parametrize("dtype", [torch.float32, torch.uint8])
def test_encode_loader_matches_reference_loop(dtype):
    from inference import encode as E
    O, spec, m, sd = _setup()
    loader = _loader(4, 6, 3, dtype)
    Z, LV, labels, paths = E.encode_loader(m, loader, torch.device("cuda"))
    assert Z.shape == (21, 8) and LV.shape == (21, 8) and Z.dtype == np.float32
    assert labels == [l for b in loader for l in b["label"]] and paths == [p for b in loader for p in b["path"]]
    with torch.no_grad():
        ref = [O.encode(sd, spec, b["image"].float().div(255) if dtype == torch.uint8 else b["image"]) for b in loader]
    assert rel_err(torch.from_numpy(Z), torch.cat([r[0] for r in ref])) <= 1e-5
    assert rel_err(torch.from_numpy(LV), torch.cat([r[1] for r in ref])) <= 1e-5
    Z0, LV0, l0, p0 = E.encode_loader(m, [], torch.device("cuda"))       # empty loader
    assert Z0.shape == (0, 8) and l0 == [] and p0 == []


def test_write_embeddings_files(tmp_path, monkeypatch):
    from inference import encode as E
    monkeypatch.setattr(E, "get_config", lambda: SimpleNamespace(paths=SimpleNamespace(tables_dir=str(tmp_path / "tables"))))
    Z = np.arange(12, dtype=np.float32).reshape(3, 4) / 7
    LV = -Z
    csv_path = E.write_embeddings(Z, LV, ["a", "b", "c"], ["p0", "p1", "p2"], "train_latents")
    assert os.path.basename(csv_path) == "train_latents_embeddings.csv"
    assert np.array_equal(np.load(tmp_path / "tables" / "train_latents_mu.npy"), Z)
    assert np.array_equal(np.load(tmp_path / "tables" / "train_latents_logvar.npy"), LV)
    rows = list(csv.reader(open(csv_path)))
    assert rows[0] == ["path", "label", "z0", "z1", "z2", "z3"]
    assert rows[2][:2] == ["p1", "b"] and [np.float32(v) for v in rows[2][2:]] == list(Z[1])


def test_batched_edit_and_interpolation_match_reference_loop(tmp_path):
    from inference import generate as G
    O, spec, m, sd = _setup(seed=5)
    g = torch.Generator().manual_seed(2)
    batch = torch.rand(3, 1, 32, 32, generator=g)
    steps, span, dim = 7, 3.0, 2
    grid = G.edit_tumor_factor(m, batch, torch.device("cuda"), dim, steps, span, str(tmp_path))
    assert os.path.getsize(tmp_path / f"edit_dim{dim}.png") > 0
    with torch.no_grad():
        mu, _ = O.encode(sd, spec, batch)
        ref = []
        for v in torch.linspace(-span, span, steps):          # reference: src/inference/generate.py:43-47
            z = mu[:1].clone()
            z[:, dim] = v
            ref.append(O.decode(sd, spec, z))
        ref = torch.cat(ref, dim=0)
    assert grid.shape == ref.shape and rel_err(grid, ref) <= 1e-5

    a, b = batch[:1], batch[1:2]
    seq = G.interpolate(m, a, b, torch.device("cuda"), steps, str(tmp_path))
    assert os.path.getsize(tmp_path / "interpolation.png") > 0
    with torch.no_grad():
        mu, _ = O.encode(sd, spec, torch.cat([a, b], dim=0))
        ref = torch.cat([O.decode(sd, spec, (1 - al) * mu[0:1] + al * mu[1:2]) for al in torch.linspace(0, 1, steps)], dim=0)
    assert rel_err(seq, ref) <= 1e-5


def test_sample_random_is_seeded_and_writes_grid(tmp_path):
    from inference import generate as G
    O, spec, m, sd = _setup(seed=7)
    a = G.sample_random(m, 9, str(tmp_path), seed=123, filename="samples.png")
    b = G.sample_random(m, 9, str(tmp_path), seed=123, filename="samples2.png")
    assert a.shape == (9, 1, 32, 32) and torch.equal(a, b)
    assert float(a.min()) > 0.0 and float(a.max()) < 1.0
    assert os.path.getsize(tmp_path / "samples.png") > 0
    torch.manual_seed(123)                                     # same draw as model.sample_prior on this device
    z = torch.randn(9, m.latent_dim, device="cuda")
    with torch.no_grad():
        ref = O.decode(sd, spec, z.cpu())
    assert rel_err(a, ref) <= 1e-5
