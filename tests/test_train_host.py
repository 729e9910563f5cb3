"""Host-side logic of the training loop (CPU): schedulers, sharded checkpoint layout, and the data-parallel
gradient bucket all-reduce of ``TrainEngine`` on two gloo ranks."""
import math
import os
import sys

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

import cfgs

REF_SRC = "/root/reference/src"


def test_beta_scheduler_matches_reference_semantics():
    from training.schedulers import BetaScheduler
    lin = BetaScheduler({"beta_schedule": {"type": "linear", "start_beta": 0.05, "end_beta": 0.15, "warmup_epochs": 25}}, 50)
    assert lin.value(0) == pytest.approx(0.05) and lin.value(25) == pytest.approx(0.15) and lin.value(40) == pytest.approx(0.15)
    assert lin.value(10) == pytest.approx(0.05 + 0.1 * 10 / 25)
    const = BetaScheduler({"beta_schedule": {"type": "constant", "end_beta": 4.0}}, 10)
    assert const.value(3) == 4.0
    cos = BetaScheduler({"beta_schedule": {"type": "cosine", "start": 0.0, "end": 2.0}}, 11)
    assert cos.value(0) == pytest.approx(0.0) and cos.value(10) == pytest.approx(2.0) and cos.value(5) == pytest.approx(1.0)
    cyc = BetaScheduler({"beta_schedule": {"type": "cyclical", "start_beta": 0.0, "end_beta": 1.0, "cycle_length": 4}}, 20)
    assert [cyc.value(e) for e in range(5)] == pytest.approx([0.0, 0.25, 0.5, 0.75, 0.0])
    fallback = BetaScheduler({"model": {"beta": 3.0}}, 5)          # no beta_schedule section -> model.beta
    assert fallback.value(2) == 3.0
    # attribute-style config (what get_config() returns)
    from utils.brain_tumor_utils.config_parser import freeze
    shipped = BetaScheduler(freeze(cfgs.shipped("beta_vae_se_debug")), 2)
    assert shipped.value(1) == pytest.approx(0.05 + 0.1 / 25)


def test_capacity_scheduler_matches_reference_semantics():
    from training.schedulers import CapacityScheduler
    cs = CapacityScheduler({"loss": {"capacity_schedule": {"enabled": True, "C_start": 30.0, "C_end": 128.0,
                                                            "warmup_epochs": 30}}}, 100)
    assert cs.value(0) == pytest.approx(30.0) and cs.value(30) == pytest.approx(128.0) and cs.value(99) == pytest.approx(128.0)
    assert cs.value(15) == pytest.approx(30.0 + 0.5 * 98.0) and cs.value(-3) == pytest.approx(30.0)
    off = CapacityScheduler({"loss": {"capacity_schedule": {"enabled": False}}}, 10)
    assert off.enabled is False and off.value(3) is None


@pytest.mark.skipif(not os.path.isdir(REF_SRC), reason="reference tree not present (GPU box)")
def test_schedulers_equal_the_reference_on_the_shipped_configs():
    import importlib.util
    spec = importlib.util.spec_from_file_location("ref_schedulers", os.path.join(REF_SRC, "training", "schedulers.py"))
    ref = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(ref)
    from training import schedulers as ours
    for name in ("beta_vae_se", "beta_vae_se_debug", "overfit_capacity"):
        cfg = cfgs.shipped(name)
        for total in (2, 100):
            for e in range(0, 60, 7):
                assert ours.BetaScheduler(cfg, total).value(e) == pytest.approx(ref.BetaScheduler(cfg, total).value(e))
                assert ours.CapacityScheduler(cfg, total).value(e) == ref.CapacityScheduler(cfg, total).value(e)


def _state():
    g = torch.Generator().manual_seed(0)
    return {f"layer{i}.{kind}": torch.randn(3, i + 1, generator=g) for i in range(5) for kind in ("weight", "bias")}


def test_sharded_checkpoint_layout_and_round_trip(tmp_path):
    from utils.brain_tumor_utils.io import load_sharded_checkpoint, save_sharded_checkpoint
    state = _state()
    base = str(tmp_path / "models" / "run_latest.pt")
    open_base = tmp_path / "models"
    open_base.mkdir()
    (open_base / "run_latest.pt").write_bytes(b"stale")          # a stale un-sharded file is removed
    paths = save_sharded_checkpoint(base, {"epoch": 3, "total_steps": 42, "model_state": state,
                                           "optim_state": {"state": {}, "param_groups": []}, "val_total": 1.5}, num_shards=2)
    assert [os.path.basename(p) for p in paths] == ["run_latest_shard0.pt", "run_latest_shard1.pt"]
    assert not os.path.exists(base)
    keys = sorted(state)
    for sid, p in enumerate(paths):
        part = torch.load(p, weights_only=False)
        assert sorted(part["model_state"]) == keys[sid::2]          # round-robin over the sorted keys
        assert part["shard_id"] == sid and part["num_shards"] == 2
        assert part["epoch"] == 3 and part["total_steps"] == 42 and part["val_total"] == 1.5 and "optim_state" in part
    back = load_sharded_checkpoint(base)
    assert sorted(back["model_state"]) == keys and back["epoch"] == 3
    for k in keys:
        assert torch.equal(back["model_state"][k], state[k])
    three = save_sharded_checkpoint(str(tmp_path / "m" / "x.pt"), {"model_state": state}, num_shards=3)
    assert len(three) == 3 and sorted(load_sharded_checkpoint(str(tmp_path / "m" / "x.pt"))["model_state"]) == keys
    with pytest.raises(ValueError):
        save_sharded_checkpoint(base, {"epoch": 1})
    with pytest.raises(FileNotFoundError):
        load_sharded_checkpoint(str(tmp_path / "nothing.pt"))


@pytest.mark.skipif(not os.path.isdir(REF_SRC), reason="reference tree not present (GPU box)")
def test_sharded_checkpoints_interoperate_with_the_reference(tmp_path):
    """Files written here load with the reference's loader and vice versa (same layout on disk)."""
    import importlib.util
    import types
    pkg = types.ModuleType("refbt")
    pkg.__path__ = [os.path.join(REF_SRC, "utils", "brain_tumor_utils")]
    sys.modules["refbt"] = pkg
    cp = types.ModuleType("refbt.config_parser")         # io.py imports get_config at module level only
    cp.get_config = lambda: None
    sys.modules["refbt.config_parser"] = cp
    spec = importlib.util.spec_from_file_location("refbt.io", os.path.join(pkg.__path__[0], "io.py"))
    ref_io = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(ref_io)
    from utils.brain_tumor_utils import io as ours
    state = _state()
    a = str(tmp_path / "a" / "ck.pt")
    ours.save_sharded_checkpoint(a, {"epoch": 7, "model_state": state}, num_shards=2)
    got = ref_io.load_sharded_checkpoint(a)
    assert got["epoch"] == 7 and all(torch.equal(got["model_state"][k], state[k]) for k in state)
    b = str(tmp_path / "b" / "ck.pt")
    ref_io.save_sharded_checkpoint(b, {"epoch": 9, "model_state": state}, num_shards=2)
    got = ours.load_sharded_checkpoint(b)
    assert got["epoch"] == 9 and all(torch.equal(got["model_state"][k], state[k]) for k in state)
    assert sorted(os.listdir(tmp_path / "a")) == sorted(os.listdir(tmp_path / "b"))


# ------------------------------------------------------------------------------------------------
# data parallel: gradient buckets over gloo, world_size 2
# ------------------------------------------------------------------------------------------------
def _dp_worker(rank, world, port, out, mode):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from bvae_rt.engine import TrainEngine, backward_ready_buckets
        from models.beta_vae import BetaVAE
        from utils.brain_tumor_utils.config_parser import freeze
        torch.manual_seed(0)
        model = BetaVAE(freeze(cfgs.small()))                      # parameters only; no kernel runs on the CPU
        names = [n for n, _ in model.named_parameters()]
        buckets = backward_ready_buckets(model)
        assert sorted(n for b in buckets for n in b) == sorted(names) and len(buckets) == 4
        assert buckets[0][0].startswith("final_conv") and buckets[-1][-1].startswith("encoder.0")
        eng = TrainEngine(model, lr=1e-3, clip=3.0, dp_mode=mode)   # flat buffers + hooks + bucket plan (CPU tensors)
        assert eng.world == world
        # the GroupNorm / SE parameter gradients are reduced at the END of backward only when no bucket leaves earlier
        from bvae_rt import ops
        assert ops._GN_DEFER["enabled"] == (mode != "buckets")
        # parameters now live in the flat buffer; every gradient slot is a view of flat_g
        for n, p in model.named_parameters():
            o, s, view = eng.slots[n]
            assert p.data_ptr() == eng.flat_p[o:o + s].data_ptr() and view.shape == p.shape
        eng.zero_grad()
        # d/dp sum((rank+1) * c_n * p) = (rank+1) * c_n : hooks fire per parameter, buckets reduce as they complete
        coef = {n: float(i + 1) for i, n in enumerate(names)}
        loss = sum(((rank + 1) * coef[n] * p).sum() for n, p in model.named_parameters())
        loss.backward()
        # "buckets": four all-reduces were launched from the hooks; "single": none yet, finish_reduce() issues the one
        assert all(c == 0 for c in eng._pending) and len(eng._works) == (4 if mode == "buckets" else 0)
        eng.finish_reduce()
        total = sum(r + 1 for r in range(world))
        for n, p in model.named_parameters():
            o, s, view = eng.slots[n]
            assert torch.allclose(view, torch.full_like(view, total * coef[n])), n      # SUM over ranks (1/world is folded into Adam)
            assert p.grad.data_ptr() == view.data_ptr()                                  # landed in the bucket, no gather copy
        eng.close()
        assert ops._GN_DEFER["enabled"]                               # restored for plain autograd use
        out[rank] = "ok"
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("mode", ["single", "buckets"])
def test_gradient_all_reduce_over_two_gloo_ranks(mode, monkeypatch):
    monkeypatch.delenv("BVAE_DP_MODE", raising=False)
    ctx = mp.get_context("spawn")
    out = ctx.Manager().dict()
    port = 29500 + (os.getpid() % 400) + (7 if mode == "single" else 0)
    procs = [ctx.Process(target=_dp_worker, args=(r, 2, port, out, mode)) for r in range(2)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(180)
    assert all(p.exitcode == 0 for p in procs), [p.exitcode for p in procs]
    assert dict(out) == {0: "ok", 1: "ok"}


def test_inference_latent_sequences_match_reference_loops():
    """generate.py's batched latent matrices are the rows the reference builds one Python iteration at a time
    (reference: src/inference/generate.py:41-46, 55-59)."""
    from inference import generate as G
    g = torch.Generator().manual_seed(1)
    mu = torch.randn(2, 8, generator=g)
    z = G.edit_latents(mu[:1], 3, 5, 3.0)
    for i, v in enumerate(torch.linspace(-3.0, 3.0, 5)):
        ref = mu[:1].clone()
        ref[:, 3] = v
        assert torch.equal(z[i:i + 1], ref)
    zi = G.interpolation_latents(mu[0:1], mu[1:2], 6)
    for i, a in enumerate(torch.linspace(0, 1, 6)):
        assert torch.equal(zi[i:i + 1], (1 - a) * mu[0:1] + a * mu[1:2])


def test_train_loop_stages_batches_one_ahead():
    """train._staged_images hands batch i to the step only after batch i+1's host->device copy was started
    (GraphedTrainStep.prefetch), keeps the loader's order, and passes through what it cannot stage."""
    from training.train import _staged_images

    class FakeStep:
        def __init__(self):
            self.x = torch.zeros(4, 1, 8, 8)
            self.log = []

        def prefetch(self, t, augment=None):
            assert augment is None
            self.log.append(("prefetch", int(t[0, 0, 0, 0])))
            return ("staged", int(t[0, 0, 0, 0]))

    step = FakeStep()
    batches = [{"image": torch.full((4, 1, 8, 8), float(i))} for i in range(4)]
    batches.append({"image": torch.full((3, 1, 8, 8), 4.0)})          # ragged tail: shape differs -> passed through
    batches.append({"image": [torch.full((1, 8, 8), 5.0)] * 4})       # list of samples -> stacked, then staged
    seen = []
    for item in _staged_images(batches, step):
        step.log.append(("step", item[1] if isinstance(item, tuple) else int(item[0, 0, 0, 0])))
        seen.append(item)
    assert [s[1] if isinstance(s, tuple) else int(s[0, 0, 0, 0]) for s in seen] == [0, 1, 2, 3, 4, 5]
    assert isinstance(seen[4], torch.Tensor) and seen[4].shape[0] == 3
    # every step i is preceded by the prefetch of batch i+1 (when that batch can be staged)
    for i in (0, 1, 2):
        assert step.log.index(("prefetch", i + 1)) < step.log.index(("step", i))
    assert step.log.index(("prefetch", 5)) < step.log.index(("step", 4))
    # without a graphed step the tensors come through untouched
    plain = list(_staged_images(batches[:2], None))
    assert all(isinstance(t, torch.Tensor) for t in plain) and len(plain) == 2
