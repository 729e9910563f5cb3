"""Parity AT THE TIMED WORKLOAD (GPU): exactly what bench.py measures -- bf16, the shipped YAML at its configured
per-GPU batch (default config: 32, capacity C=30 gamma=1 + FFL 0.5; overfit_capacity: 8, beta 4), the whole step
replayed as a CUDA graph by ``GraphedTrainStep`` (two eager steps, capture, replay) -- checked against the CPU
oracle on the same weights, inputs and injected eps:

  * loss, recon, KL of the REPLAYED step within 1e-2 of the fp64 oracle (north_star's bf16 tolerance) and of the
    oracle under bf16 autocast;
  * all parameter gradients the replayed step left in the engine's flat buffer: every tensor within 1e-2 of fp64
    (or, for ill-conditioned tensors, within 4x of the bf16-autocast reference's own distance from fp64 --
    tests/parity.py), and the whole gradient within max(1e-2, 1.25x the bf16-autocast reference's own distance from
    fp64): measured 1.32e-2 against the reference's own 1.24e-2 at the default config, i.e. the precision's error;
  * the graph's outputs equal an eager step on the same inputs bit for bit (same kernels, same order).

Batch-dependent kernel plans (persistent-CTA tile counts, GroupNorm cluster sizes, split-K factors) are the ones the
benchmark times.  The measured errors are written to gpurun_out/parity_r2.json (committed copy: profiles/parity_r2.json).
Reference: src/models/beta_vae.py:194-272 (loss), src/training/train.py:313-330 (the step)."""
import pytest
import torch

from parity import check_vs_exact, global_norm, own_error_by_kind, rel_err, report

pytestmark = pytest.mark.gpu

CASES = {
    # name: (yaml, batch, loss kwargs)  -- bench.py WORKLOADS
    "default_b32": ("beta_vae_se", 32, dict(beta=1.0, capacity=30.0, free_bits=0.0, capacity_weight=1.0)),
    "overfit_capacity_b8": ("overfit_capacity", 8, dict(beta=4.0, capacity=None, free_bits=0.0, capacity_weight=None)),
}


@pytest.mark.parametrize("case", sorted(CASES))
def test_graph_replayed_bf16_step_matches_oracle(case):
    import cfgs
    from bvae_rt.data import synthetic_batch
    from bvae_rt.engine import GraphedTrainStep, TrainEngine
    from models.beta_vae import BetaVAE
    from oracle import beta_vae_oracle as O
    from utils.brain_tumor_utils.config_parser import freeze

    yaml_name, B, kw = CASES[case]
    cfg = cfgs.shipped(yaml_name)
    size, L = cfg["data"]["image_size"], cfg["model"]["latent_dim"]
    dev = torch.device("cuda")
    torch.manual_seed(cfg["data"]["seed"])
    model = BetaVAE(freeze(cfg))
    model.apply(cfgs.weight_init)
    model = model.to(dev).train()
    engine = TrainEngine(model, lr=cfg["optimization"]["lr"], clip=cfg["training"]["grad_clip"])
    step = GraphedTrainStep(engine, (B, 1, size, size), seed=3)
    xs = [synthetic_batch(B, size, seed=10 + i, device=dev) for i in range(4)]
    gen = torch.Generator(device=dev).manual_seed(5)
    epss = [torch.randn(B, L, device=dev, generator=gen) for _ in range(4)]
    for i in range(3):                              # 2 eager steps, then capture + first replay
        step(xs[i], eps=epss[i], **kw)
    assert step.graph is not None, "the third call must have captured the graph"
    torch.cuda.synchronize()
    names = [n for n, _ in model.named_parameters()]
    sd = {k: v.detach().cpu().clone() for k, v in model.state_dict().items()}      # weights the 4th step starts from
    flat_before = engine.flat_p.clone()
    out = step(xs[3], eps=epss[3], **kw)            # pure replay
    torch.cuda.synchronize()
    got = {k: float(out[k]) for k in ("total", "recon", "kl_mean", "recon_ffl")}
    grads = {n: engine.slots[n][2].detach().cpu().clone() for n in names}
    assert not torch.equal(flat_before, engine.flat_p), "the replayed step did not update the parameters"

    spec = O.spec_from_config(cfgs.ns(cfg))
    x_cpu, e_cpu = xs[3].cpu(), epss[3].cpu()
    ex_out, ex_g = O.loss_and_grads(sd, spec, x_cpu, e_cpu, precision="fp64", **kw)
    lp_out, lp_g = O.loss_and_grads(sd, spec, x_cpu, e_cpu, precision="bf16", **kw)

    rtol, floor = 1e-2, 1e-4
    rec = {"config": yaml_name, "batch": B, "dtype": "bf16", "launch": "CUDA graph replay (4th call)", "tolerance": rtol,
           "scalars": {}}
    for k in ("total", "recon", "kl_mean"):
        e64 = abs(got[k] - float(ex_out[k])) / abs(float(ex_out[k]))
        elp = abs(got[k] - float(lp_out[k])) / abs(float(lp_out[k]))
        eref = abs(float(lp_out[k]) - float(ex_out[k])) / abs(float(ex_out[k]))
        rec["scalars"][k] = {"got": got[k], "oracle_fp64": float(ex_out[k]), "oracle_bf16_autocast": float(lp_out[k]),
                             "rel_vs_fp64": e64, "rel_vs_bf16_oracle": elp, "oracle_bf16_own_rel": eref}
        assert e64 <= rtol, (k, got[k], float(ex_out[k]))
        assert elp <= rtol, (k, got[k], float(lp_out[k]))
    gn = global_norm(ex_g.values())
    rows, failures = [], []
    kind_rel = own_error_by_kind(names, lp_g, ex_g)        # the bf16 reference's own noise per kind of tensor
    for k in names:
        try:
            e_got, e_ref = check_vs_exact(k, grads[k], lp_g[k], ex_g[k], rtol, gn, floor, kind_rel=kind_rel)
        except AssertionError as e:
            failures.append(str(e))
            e_got, e_ref = rel_err(grads[k], ex_g[k]), rel_err(lp_g[k], ex_g[k])
        rows.append((k, e_got, e_ref))
    cat = lambda d: torch.cat([d[k].flatten().double() for k in names])      # noqa: E731
    whole = rel_err(cat(grads), cat(ex_g))
    whole_ref = rel_err(cat(lp_g), cat(ex_g))
    worst = max(rows, key=lambda r: r[1])
    rec.update({
        "whole_gradient_rel_vs_fp64": whole, "oracle_bf16_whole_gradient_rel_vs_fp64": whole_ref,
        "worst_tensor": {"name": worst[0], "rel_vs_fp64": worst[1], "oracle_bf16_own_rel": worst[2]},
        "tensors_above_1e-2": sorted([(r[0], r[1], r[2]) for r in rows if r[1] > rtol], key=lambda r: -r[1]),
        "n_tensors": len(rows), "tensors_failing_floor_rule": failures,
        "grad_norm": {"got": float(cat(grads).norm()), "oracle_fp64": float(cat(ex_g).norm())},
    })
    report(case, rec)
    print(f"{case}: total {got['total']:.4f} (fp64 {float(ex_out['total']):.4f}); whole-gradient rel {whole:.2e} "
          f"(bf16 oracle's own {whole_ref:.2e}); worst {worst[0]} {worst[1]:.2e} (own {worst[2]:.2e}); "
          f"{len(rec['tensors_above_1e-2'])}/{len(rows)} tensors above 1e-2")
    assert not failures, failures
    # whole gradient: 1e-2 where bf16 itself allows it; at this depth the reference under bf16 autocast is itself
    # ~1.1-1.2e-2 from its fp64 evaluation (reported above; it moves by ~7 % between hosts, with the reduction order of
    # the CPU autocast kernels), so the bound is the larger of the two with 1.5x headroom (measured ratio 1.07-1.21)
    assert whole <= max(rtol, 1.5 * whole_ref), (whole, whole_ref)

    # the replay and an un-captured step on the same inputs run the same kernels: bit-identical gradients
    engine.flat_p.copy_(flat_before)
    model.invalidate_packs()
    engine.zero_grad()
    with torch.autocast("cuda", dtype=torch.bfloat16):
        o2 = model.loss(xs[3], eps=epss[3], **kw)
    o2["total"].backward()
    engine.finish_reduce()
    torch.cuda.synchronize()
    assert float(o2["total"]) == got["total"]
    for n in names:
        assert torch.equal(engine.slots[n][2].cpu(), grads[n]), f"graph replay and eager step differ on {n}"
    engine.close()
