"""CPU oracle for the Beta-VAE-SE training hot path.  TEST INFRASTRUCTURE ONLY.

This module is the *checker*, never the product: only ``tests/``,
``__graft_entry__.smoke()`` and ``bench.py``'s ``cpu_baseline`` / ``--impl
reference`` legs may import it.  The product path (``beta-vae_b200/``) never
imports anything under ``oracle/`` and raises if its CUDA library is missing.

It restates, as plain functional fp32 PyTorch over a ``state_dict``, what the
reference computes on the path named by BASELINE.json's north_star
(reference paths are relative to the upstream repo root):

  * ConvBlock / DeconvBlock forward        src/models/beta_vae.py:33-66
  * GroupNorm(1, C) + ReLU                 src/models/beta_vae.py:17-24, 8-15
  * SEBlock                                src/models/se_blocks.py:4-19
  * encode / reparameterize / decode       src/models/beta_vae.py:135-165
  * reconstruction_loss                    src/models/beta_vae.py:185-192
  * loss (KL, free bits, capacity, reg)    src/models/beta_vae.py:194-272
  * FocalFrequencyLoss                     src/training/losses.py:16-35

The arithmetic of the reference lives in third-party PyTorch (unpinned in the
reference's requirements.txt; its authors ran torch 2.7.0, this image has
2.11.0).  The reference ships no tests or golden vectors, so parity is pinned
by us: ``tests/golden/make_golden.py`` imports the *unmodified* reference from
/root/reference in the build container, runs it on seeded inputs and commits
the resulting vectors; ``tests/test_oracle_golden.py`` checks this restatement
against them (and against the survey-time scalars in SURVEY.md section 8c).

Everything that is elementwise / a reduction is written out explicitly (group
norm statistics, the {0.25, 0.75} bilinear stencil, SE gating, KL, FFL) so the
CUDA kernels have a formula-level specification; only the dense contractions
(conv2d / linear / fft2) call the library primitive the reference itself calls.
Gradients come from autograd over these formulas.
"""
from __future__ import annotations

from dataclasses import dataclass, field
from typing import Dict, Optional, Sequence, Tuple

import torch
import torch.nn.functional as F


@dataclass(frozen=True)
class PathSpec:
    """Hyper-parameters of the path; names follow the reference's YAML keys
    (configs/beta_vae_se.yaml:11-50)."""

    image_size: int = 128
    in_ch: int = 1                      # data.grayscale -> 1 else 3   (beta_vae.py:81)
    latent_dim: int = 64
    base_channels: int = 64
    num_blocks: int = 4
    se_reduction_ratio: int = 8
    use_decoder_se: bool = True
    reconstruction_loss: str = "mse"    # mse | l1 | bce
    logvar_clamp: Optional[Tuple[float, float]] = (-10.0, 5.0)
    latent_clamp: Optional[float] = None
    latent_reg_lambda: float = 0.0
    deterministic: bool = False         # model.deterministic_overfit
    use_ffl: bool = False
    ffl_weight: float = 0.0
    ffl_alpha: float = 1.0

    # ---- derived (beta_vae.py:94-125) ----
    @property
    def enc_channels(self):
        return [self.in_ch] + [self.base_channels * (2 ** i) for i in range(self.num_blocks)]

    @property
    def dec_channels(self):
        """(in, out) per decoder block: reversed encoder widths, last block keeps its width."""
        rev = list(reversed(self.enc_channels[1:]))
        out = []
        for i in range(self.num_blocks):
            out.append((rev[i], rev[i + 1] if i + 1 < len(rev) else rev[-1]))
        return out

    @property
    def enc_out_shape(self):
        s = self.image_size
        for _ in range(self.num_blocks):
            s = (s + 2 - 3) // 2 + 1     # Conv2d(3, stride=2, padding=1)
        return (self.enc_channels[-1], s, s)

    @property
    def flat_dim(self):
        c, h, w = self.enc_out_shape
        return c * h * w


def spec_from_config(cfg) -> PathSpec:
    """Build a PathSpec from an attribute-style config (same keys the reference reads)."""
    m, d = cfg.model, cfg.data
    l = getattr(cfg, "loss", None)
    if getattr(m, "encoder_norm", "layer") != "layer":
        raise NotImplementedError("oracle restates encoder_norm=layer only")
    if getattr(m, "activation", "relu") != "relu":
        raise NotImplementedError("oracle restates activation=relu only")
    if getattr(m, "encoder_pooling", "flatten") != "flatten":
        raise NotImplementedError("oracle restates encoder_pooling=flatten only")
    lc = getattr(m, "logvar_clamp", None)
    return PathSpec(
        image_size=d.image_size,
        in_ch=1 if d.grayscale else 3,
        latent_dim=m.latent_dim,
        base_channels=m.base_channels,
        num_blocks=m.num_blocks,
        se_reduction_ratio=m.se_reduction_ratio,
        use_decoder_se=m.use_decoder_se,
        reconstruction_loss=m.reconstruction_loss,
        logvar_clamp=tuple(lc) if lc else None,
        latent_clamp=getattr(m, "latent_clamp", None),
        latent_reg_lambda=getattr(m, "latent_reg_lambda", 0.0),
        deterministic=getattr(m, "deterministic_overfit", False),
        use_ffl=bool(getattr(l, "use_ffl", False)),
        ffl_weight=float(getattr(l, "ffl_weight", 0.0)),
        ffl_alpha=float(getattr(l, "ffl_alpha", 1.0)),
    )


# --------------------------------------------------------------------------------------
# layer formulas
# --------------------------------------------------------------------------------------

def group_norm_1(x: torch.Tensor, gamma: torch.Tensor, beta: torch.Tensor, eps: float = 1e-5):
    """nn.GroupNorm(1, C): one group = per-sample statistics over C*H*W, biased variance,
    per-channel affine (beta_vae.py:20-21)."""
    if x.dtype in (torch.bfloat16, torch.float16):
        x = x.float()          # autocast runs group_norm in fp32 (ATen autocast fp32 list)
    n = x.shape[0]
    flat = x.reshape(n, -1)
    mean = flat.mean(dim=1)
    var = (flat - mean[:, None]).pow(2).mean(dim=1)
    rstd = torch.rsqrt(var + eps)
    xn = (x - mean[:, None, None, None]) * rstd[:, None, None, None]
    return xn * gamma[None, :, None, None] + beta[None, :, None, None]


def se_gate(x, w1, b1, w2, b2):
    """SEBlock.forward (se_blocks.py:15-19): squeeze = mean over H,W; excitation =
    sigmoid(W2 relu(W1 y + b1) + b2); rescale x by the per-(n,c) gate."""
    y = x.mean(dim=(2, 3))
    h = torch.relu(y @ w1.t() + b1)
    s = torch.sigmoid(h @ w2.t() + b2)
    return x * s[:, :, None, None]


def _up2_axis(x: torch.Tensor, dim: int) -> torch.Tensor:
    """One axis of bilinear x2, align_corners=False: out[2i] = .25 x[i-1] + .75 x[i],
    out[2i+1] = .75 x[i] + .25 x[i+1], indices clamped at the border."""
    n = x.shape[dim]
    idx = torch.arange(n, device=x.device)
    prev = x.index_select(dim, (idx - 1).clamp(min=0))
    nxt = x.index_select(dim, (idx + 1).clamp(max=n - 1))
    even = 0.25 * prev + 0.75 * x
    odd = 0.75 * x + 0.25 * nxt
    out = torch.stack((even, odd), dim=dim + 1)
    shape = list(x.shape)
    shape[dim] = 2 * n
    return out.reshape(shape)


def upsample2x_bilinear(x: torch.Tensor) -> torch.Tensor:
    """nn.Upsample(scale_factor=2, mode='bilinear', align_corners=False) (beta_vae.py:53);
    the stencil is separable."""
    return _up2_axis(_up2_axis(x, 2), 3)


def conv_block(x, sd, prefix, conv_key, stride, has_se):
    """conv3x3(pad 1) -> GroupNorm(1,C) -> ReLU -> SE (beta_vae.py:41-46, 61-66)."""
    x = F.conv2d(x, sd[f"{prefix}.{conv_key}.weight"], sd[f"{prefix}.{conv_key}.bias"],
                 stride=stride, padding=1)
    x = group_norm_1(x, sd[f"{prefix}.norm.weight"], sd[f"{prefix}.norm.bias"])
    x = torch.relu(x)
    if has_se:
        x = se_gate(x, sd[f"{prefix}.se.block.fc.0.weight"], sd[f"{prefix}.se.block.fc.0.bias"],
                    sd[f"{prefix}.se.block.fc.2.weight"], sd[f"{prefix}.se.block.fc.2.bias"])
    return x


# --------------------------------------------------------------------------------------
# model formulas
# --------------------------------------------------------------------------------------

def encode(sd: Dict[str, torch.Tensor], spec: PathSpec, x: torch.Tensor):
    """BetaVAE.encode (beta_vae.py:135-145). Encoder SE is always on (beta_vae.py:99)."""
    h = x
    for i in range(spec.num_blocks):
        h = conv_block(h, sd, f"encoder.{i}", "conv", 2, True)
    h = h.reshape(h.shape[0], -1)               # NCHW flatten: feature = c*H*W + y*W + x
    mu = h @ sd["fc_mu.weight"].t() + sd["fc_mu.bias"]
    logvar = h @ sd["fc_logvar.weight"].t() + sd["fc_logvar.bias"]
    if mu.dtype in (torch.bfloat16, torch.float16):   # under autocast exp/pow/sum run in fp32
        mu, logvar = mu.float(), logvar.float()
    lo, hi = spec.logvar_clamp if spec.logvar_clamp else (-10.0, 10.0)
    logvar = torch.clamp(logvar, lo, hi)        # gradient is zero outside [lo, hi]
    return mu, logvar


def reparameterize(mu, logvar, eps):
    """BetaVAE.reparameterize (beta_vae.py:147-150) with eps supplied by the caller
    (the reference draws it internally with randn_like)."""
    return mu + eps * torch.exp(0.5 * logvar)


def decode(sd, spec: PathSpec, z):
    """BetaVAE.decode (beta_vae.py:152-165)."""
    if spec.latent_clamp is not None:
        z = torch.clamp(z, -spec.latent_clamp, spec.latent_clamp)
    h = z @ sd["fc_dec.weight"].t() + sd["fc_dec.bias"]
    h = h.reshape(z.shape[0], *spec.enc_out_shape)
    for i in range(spec.num_blocks):
        h = upsample2x_bilinear(h)
        h = conv_block(h, sd, f"decoder_blocks.{i}", "up.1", 1, spec.use_decoder_se)
    logit = F.conv2d(h, sd["final_conv.weight"], sd["final_conv.bias"], stride=1, padding=1)
    return torch.sigmoid(logit)


def forward(sd, spec: PathSpec, x, eps=None, deterministic=None):
    """BetaVAE.forward (beta_vae.py:167-183)."""
    if deterministic is None:
        deterministic = spec.deterministic
    mu, logvar = encode(sd, spec, x)
    z = mu if deterministic else reparameterize(mu, logvar, eps)
    return decode(sd, spec, z), mu, logvar, z


def reconstruction_loss(spec: PathSpec, recon, x):
    """Sum over all elements divided by batch size (beta_vae.py:185-192)."""
    b = x.shape[0]
    if spec.reconstruction_loss == "mse":
        return (recon - x).pow(2).sum() / b
    if spec.reconstruction_loss == "l1":
        return (recon - x).abs().sum() / b
    if spec.reconstruction_loss == "bce":
        # F.binary_cross_entropy clamps each log term at -100
        lp = torch.log(recon).clamp(min=-100.0)
        l1p = torch.log(1.0 - recon).clamp(min=-100.0)
        return -(x * lp + (1.0 - x) * l1p).sum() / b
    raise ValueError("invalid reconstruction_loss")


def focal_frequency_loss(pred, target, alpha: float = 1.0, eps: float = 1e-8):
    """FocalFrequencyLoss.forward (losses.py:16-35): the weight is NOT detached."""
    up = (lambda t: t) if pred.dtype == torch.float64 else (lambda t: t.float())
    pf = torch.fft.fft2(up(pred), norm="ortho")
    tf = torch.fft.fft2(up(target), norm="ortho")
    d = pf - tf
    dist = d.real.pow(2) + d.imag.pow(2)
    denom = dist.mean(dim=(0, 2, 3), keepdim=True) + eps
    weight = (dist / denom).clamp_min(eps).pow(alpha)
    return (weight * dist).mean()


def loss(sd, spec: PathSpec, x, eps, beta=1.0, capacity=None, free_bits=0.0, capacity_weight=None):
    """BetaVAE.loss (beta_vae.py:194-272). LPIPS is out of scope (third-party network)."""
    recon, mu, logvar, z = forward(sd, spec, x, eps)
    base = reconstruction_loss(spec, recon, x)
    ff = torch.zeros((), dtype=x.dtype)
    if spec.use_ffl and spec.ffl_weight > 0:
        ff = focal_frequency_loss(recon, x, spec.ffl_alpha) * spec.ffl_weight
    rec = base + ff
    use_capacity = (capacity is not None) and (capacity_weight is not None)
    if spec.deterministic:
        kl_per_dim = torch.zeros(spec.latent_dim)
        kl_mean = torch.zeros(())
        kl_eff = torch.zeros(())
    else:
        kl_elem = -0.5 * (1 + logvar - mu.pow(2) - logvar.exp())
        kl_per_dim = kl_elem.mean(dim=0)
        kl_mean = kl_elem.sum(dim=1).mean()
        if free_bits > 0 and not use_capacity:
            kl_eff = torch.clamp(kl_per_dim, min=free_bits).sum()
        else:
            kl_eff = kl_per_dim.sum()
    reg = 0.0
    if spec.latent_reg_lambda > 0:
        reg = spec.latent_reg_lambda * mu.pow(2).mean()
    if spec.deterministic:
        total = rec + reg
    elif use_capacity:
        total = rec + capacity_weight * (kl_mean - capacity).abs() + reg
    else:
        total = rec + beta * kl_eff + reg
    return {
        "total": total, "recon": rec, "recon_base": base, "recon_ffl": ff,
        "kl_mean": kl_mean, "kl_per_dim": kl_per_dim, "kl_effective": kl_eff,
        "recon_img": recon, "z": z, "mu": mu, "logvar": logvar,
        "mode": "capacity" if use_capacity else "beta",
    }


def loss_and_grads(sd, spec: PathSpec, x, eps, precision="fp32", **kw):
    """fwd + bwd: returns (loss dict with detached tensors, {key: grad}).
    precision: "fp32" (the reference with mixed_precision: false), "fp64" (exact answer),
    "bf16" (the reference's arithmetic under torch.autocast(cpu, bfloat16): conv / linear in
    bf16, normalisation / loss math in fp32 -- used to size the bf16 tolerance)."""
    dt = torch.float64 if precision == "fp64" else torch.float32
    leaf = {k: v.detach().clone().to(dt).requires_grad_(True) for k, v in sd.items()}
    xe, ee = x.to(dt), None if eps is None else eps.to(dt)
    if precision == "bf16":
        with torch.autocast("cpu", dtype=torch.bfloat16):
            out = loss(leaf, spec, xe, ee, **kw)
    else:
        out = loss(leaf, spec, xe, ee, **kw)
    out["total"].backward()
    grads = {k: (v.grad if v.grad is not None else torch.zeros_like(v)) for k, v in leaf.items()}
    det = {k: (v.detach() if torch.is_tensor(v) else v) for k, v in out.items()}
    return det, grads


# --------------------------------------------------------------------------------------
# parameter construction (train.py:217-221 weight_init, beta_vae.py:94-127 module order)
# --------------------------------------------------------------------------------------

def param_shapes(spec: PathSpec):
    """state_dict keys -> shapes in the reference's ``parameters()`` order."""
    out = []
    ch = spec.enc_channels
    def se(prefix, c):
        r = max(1, c // spec.se_reduction_ratio)
        return [(f"{prefix}.se.block.fc.0.weight", (r, c)), (f"{prefix}.se.block.fc.0.bias", (r,)),
                (f"{prefix}.se.block.fc.2.weight", (c, r)), (f"{prefix}.se.block.fc.2.bias", (c,))]
    for i in range(spec.num_blocks):
        p = f"encoder.{i}"
        out += [(f"{p}.conv.weight", (ch[i + 1], ch[i], 3, 3)), (f"{p}.conv.bias", (ch[i + 1],)),
                (f"{p}.norm.weight", (ch[i + 1],)), (f"{p}.norm.bias", (ch[i + 1],))]
        out += se(p, ch[i + 1])
    out += [("fc_mu.weight", (spec.latent_dim, spec.flat_dim)), ("fc_mu.bias", (spec.latent_dim,)),
            ("fc_logvar.weight", (spec.latent_dim, spec.flat_dim)), ("fc_logvar.bias", (spec.latent_dim,)),
            ("fc_dec.weight", (spec.flat_dim, spec.latent_dim)), ("fc_dec.bias", (spec.flat_dim,))]
    for i, (ci, co) in enumerate(spec.dec_channels):
        p = f"decoder_blocks.{i}"
        out += [(f"{p}.up.1.weight", (co, ci, 3, 3)), (f"{p}.up.1.bias", (co,)),
                (f"{p}.norm.weight", (co,)), (f"{p}.norm.bias", (co,))]
        if spec.use_decoder_se:
            out += se(p, co)
    out += [("final_conv.weight", (spec.in_ch, spec.dec_channels[-1][1], 3, 3)),
            ("final_conv.bias", (spec.in_ch,))]
    return out


def init_params(spec: PathSpec, seed: int) -> Dict[str, torch.Tensor]:
    """Kaiming-normal(relu) weights, zero biases, GroupNorm gamma=1 beta=0 (train.py:217-221).
    Statistically the reference's init; NOT bit-identical to ``BetaVAE().apply(weight_init)``
    (module construction consumes RNG first) -- the product model class reproduces that."""
    g = torch.Generator().manual_seed(seed)
    sd = {}
    for k, shp in param_shapes(spec):
        if ".norm.weight" in k:
            sd[k] = torch.ones(shp)
        elif k.endswith(".bias"):
            sd[k] = torch.zeros(shp)
        else:
            fan_in = 1
            for s in shp[1:]:
                fan_in *= s
            sd[k] = torch.randn(shp, generator=g) * (2.0 / fan_in) ** 0.5
    return sd


def synthetic_batch(n: int, size: int, seed: int = 0) -> torch.Tensor:
    """Synthetic MRI-shaped slices with the class patterns of
    scripts/generate_demo_data.py:27-47 (glioma disc, meningioma bands, pituitary
    cross-hatch, notumor noise), quantised to 1/255; [n,1,size,size] fp32 in [0,1]."""
    g = torch.Generator().manual_seed(seed)
    yy, xx = torch.meshgrid(torch.arange(size), torch.arange(size), indexing="ij")
    out = torch.empty(n, 1, size, size)
    for i in range(n):
        arr = (0.25 + 0.08 * torch.randn(size, size, generator=g)).clamp(0, 1)
        cls = i % 4
        if cls == 0:
            m = (xx - size // 2) ** 2 + (yy - size // 2) ** 2 <= (size // 4) ** 2
            arr = arr + 0.35 * m
        elif cls == 1:
            m = (yy % max(1, size // 8)) < (size // 16)
            arr = arr + 0.25 * m
        elif cls == 2:
            p = max(1, size // 6)
            m = (((xx + yy) % p) < (size // 16)) | (((xx - yy) % p) < (size // 16))
            arr = arr + 0.25 * m
        else:
            arr = arr + 0.02 * torch.randn(size, size, generator=g)
        out[i, 0] = torch.floor(arr.clamp(0, 1) * 255.0) / 255.0
    return out
