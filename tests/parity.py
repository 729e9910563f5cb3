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
