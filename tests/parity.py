"""Shared parity metric.

Per-tensor relative L2 error  ||got - ref|| / ||ref||  must be <= rtol
(1e-5 for fp32, 1e-2 for bf16: BASELINE.json north_star).  Gradient tensors whose norm
is a vanishing fraction of the global gradient norm (SE excitation weights at batch 1,
for instance) carry fp32 round-off of the *reference itself* above 1e-5 of their own
norm -- measured against an fp64 evaluation the reference is 1.05e-5 off on such a
tensor -- so each gradient tensor also gets an absolute floor of  floor * ||all grads||
(floor = 1e-8 in fp32, 1e-4 in bf16)."""
import torch


def rel_err(got: torch.Tensor, ref: torch.Tensor) -> float:
    got, ref = got.detach().double().cpu(), ref.detach().double().cpu()
    d = (got - ref).norm()
    n = ref.norm()
    return float(d / n) if n > 0 else float(d)


def global_norm(tensors) -> float:
    return float(torch.sqrt(sum(t.detach().double().pow(2).sum() for t in tensors)))


def check(name, got, ref, rtol, gnorm=0.0, floor=0.0):
    got_d, ref_d = got.detach().double().cpu(), ref.detach().double().cpu()
    assert got_d.shape == ref_d.shape, (name, got_d.shape, ref_d.shape)
    assert torch.isfinite(got_d).all(), f"{name}: non-finite values"
    err = float((got_d - ref_d).norm())
    bound = rtol * float(ref_d.norm()) + floor * gnorm
    assert err <= bound or err == 0.0, f"{name}: |err|={err:.3e} > bound {bound:.3e} (rel {rel_err(got, ref):.3e})"
    return rel_err(got, ref)


def tensor_kind(name: str) -> str:
    """'encoder.2.se.block.fc.0.weight' / 'decoder_blocks.0.se.block.fc.0.weight' -> 'se.block.fc.0.weight': the same
    parameter of every encoder / decoder block."""
    parts = name.split(".")
    if parts[0] in ("encoder", "decoder_blocks") and len(parts) > 2 and parts[1].isdigit():
        return ".".join(parts[2:])
    return name


def own_error_by_kind(names, ref_lp, exact):
    """max over the tensors of a kind of the low-precision reference's own relative error vs ``exact``.

    The reference's error on ONE small tensor is a single draw of rounding noise: on a 32-element SE bias under bf16 it
    moves by 2x between two CPU hosts (thread count changes the reduction order of the autocast reference) and ours
    moves as much when a kernel's summation order changes, so the ratio of the two is heavy-tailed.  The largest error
    the reference makes on any block's tensor of the same kind is the steadier yardstick for "the noise this kind of
    tensor carries at this precision"."""
    out = {}
    for k in names:
        e = exact[k].detach().double().cpu()
        n = float(e.norm())
        r = float((ref_lp[k].detach().double().cpu() - e).norm()) / (n if n > 0 else 1.0)
        kind = tensor_kind(k)
        out[kind] = max(out.get(kind, 0.0), r)
    return out


def check_vs_exact(name, got, ref_lp, exact, rtol, gnorm=0.0, floor=0.0, slack=4.0, kind_rel=None):
    """Parity for tensors whose low-precision *reference* value is itself noisy.

    ``exact`` is the same computation in fp64 (reference module in double / oracle in double),
    ``ref_lp`` the reference at the precision under test (fp32, or bf16 autocast).  ``got``
    passes when it is within ``rtol`` of the exact answer (plus the floor above), or -- for
    ill-conditioned tensors where even the reference is further than that from exact, e.g. the
    first conv's weight gradient (sum of x*dy with sum(dy) ~ 0 after GroupNorm) -- when it is no
    further from exact than ``slack`` x the reference's own error (slack 4: the CPU reference
    accumulates its reductions in double -- ATen's acc_type<float, cpu> -- while CUDA code,
    the reference's included, accumulates in float; and two independent roundings of the same
    quantity differ by up to ~2x their individual error; measured worst case over the shipped configs: 3.5x, on
    an SE excitation weight whose fp32 reference gradient is itself 1.4e-5 off).  ``kind_rel`` (bf16 tests): the
    reference's own relative error is taken as the larger of this tensor's and the worst over the blocks' tensors of
    the same kind (``own_error_by_kind``).  Returns (err_got, err_ref) as fractions of ||exact||."""
    g, r, e = (t.detach().double().cpu() for t in (got, ref_lp, exact))
    assert g.shape == e.shape, (name, g.shape, e.shape)
    assert torch.isfinite(g).all(), f"{name}: non-finite values"
    n = float(e.norm())
    e_got, e_ref = float((g - e).norm()), float((r - e).norm())
    e_yard = e_ref if kind_rel is None else max(e_ref, kind_rel.get(tensor_kind(name), 0.0) * n)
    bound = max(rtol * n + floor * gnorm, slack * e_yard)
    assert e_got <= bound, (f"{name}: |got-exact|={e_got:.3e} > max(rtol*|exact|+floor={rtol * n + floor * gnorm:.3e}, "
                            f"{slack}*|ref-exact|={slack * e_yard:.3e})  (|exact|={n:.3e})")
    d = n if n > 0 else 1.0
    return e_got / d, e_ref / d


def report_path(name="parity_r2.json"):
    """Where the GPU parity tests drop their measured errors: gpurun_out/ of the tree they run in (it is merged back
    to the build container; the committed copy lives under profiles/)."""
    import os
    root = os.environ.get("GRAFT_REPO_ROOT") or os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    d = os.path.join(root, "gpurun_out")
    os.makedirs(d, exist_ok=True)
    return os.path.join(d, name)


def report(key, record, name="parity_r2.json"):
    """Merge one record into the JSON report (one key per test case)."""
    import json
    import os
    path = report_path(name)
    data = {}
    if os.path.exists(path):
        try:
            with open(path) as f:
                data = json.load(f)
        except Exception:  # noqa: BLE001
            data = {}
    data[key] = record
    with open(path, "w") as f:
        json.dump(data, f, indent=1, sort_keys=True)
