"""Beta-VAE with SE blocks on B200 -- drop-in for the reference's ``models.beta_vae.BetaVAE``.

Same constructor, public attributes, method signatures, ``state_dict`` keys / shapes / dtypes
and ``loss()`` dictionary as the reference class (reference: src/models/beta_vae.py:68-287),
so ``train.py``, the checkpoint code and the inference scripts run unchanged.  What differs is
what executes: every layer below dispatches to hand-written sm_100a kernels through the C ABI
in ``include/bvae_b200.h`` (no ATen conv / group_norm / upsample, no CPU path):

  encoder ConvBlock / decoder DeconvBlock -> ``ops.ConvGnSeBlock``   (beta_vae.py:33-66)
  fc_mu, fc_logvar, clamp, reparam, KL     -> ``ops.HeadsReparamKL`` (beta_vae.py:137-150, 219-253)
  latent clamp + fc_dec                    -> ``ops.FcDec``          (beta_vae.py:153-161)
  final_conv + sigmoid + recon loss        -> ``ops.FinalConvSigmoidRecon`` (beta_vae.py:163-164, 185-192)

The nn.Conv2d / nn.Linear / nn.GroupNorm sub-modules exist only to own the fp32 master
parameters under the reference's names (and to consume the RNG in the reference's order, so
``torch.manual_seed(s); BetaVAE().apply(weight_init)`` yields bit-identical weights).

Precision: ``precision="auto"`` (default) computes in bf16 storage / fp32 accumulation when
called inside ``torch.autocast("cuda")`` -- where the reference would run fp16 autocast,
SURVEY.md fact 5 -- and in exact fp32 otherwise (the reference's validation pass).
Extra, backwards-compatible hooks: ``forward/loss(..., eps=...)`` inject the reparameterisation
noise (the reference draws it internally, beta_vae.py:149).
"""
from __future__ import annotations

import torch
import torch.nn as nn

from bvae_rt import ops
from bvae_rt.lib import BF16, F32, LOSS
from training.losses import FocalFrequencyLoss, LPIPSLoss
from utils.brain_tumor_utils.config_parser import get_config

from .se_blocks import SEBlock


def _get_activation(name):
    if name == "relu":
        return nn.ReLU(inplace=True)
    if name in ("leakyrelu", "elu"):
        raise NotImplementedError(f"activation '{name}': the fused GN+ReLU+SE kernel covers relu only")
    raise ValueError("unsupported activation")


def _get_norm(name, ch):
    if name == "layer":
        return nn.GroupNorm(1, ch)
    if name in ("batch", "none"):
        raise NotImplementedError(f"encoder_norm '{name}': the fused kernel covers 'layer' (GroupNorm(1,C)) only")
    raise ValueError("unsupported norm")


class SEBlockWrapper(nn.Module):
    def __init__(self, channels, use_se, reduction):
        super().__init__()
        self.block = SEBlock(channels, reduction) if use_se else nn.Identity()

    @property
    def gate_params(self):
        return self.block.gate_params if isinstance(self.block, SEBlock) else (None, None, None, None)

    def forward(self, x):
        return self.block(x)


class _FusedBlock(nn.Module):
    """Shared forward of ConvBlock / DeconvBlock over NHWC activations."""

    stride = 1
    upsample = False

    def _conv(self) -> nn.Conv2d:
        raise NotImplementedError

    def forward(self, x, packs: ops.PackCache, key: str, dt: int, up_after: bool = False):
        """``x`` is the block's conv input (a decoder block's input arrives already upsampled); ``up_after`` makes the
        block emit the bilinear x2 upsample of its activation, i.e. the NEXT decoder block's conv input."""
        conv = self._conv()
        w1, b1, w2, b2 = self.se.gate_params
        return ops.ConvGnSeBlock.apply(x, conv.weight, conv.bias, self.norm.weight, self.norm.bias, w1, b1, w2, b2,
                                       packs, key, self.stride, bool(up_after), dt)


class ConvBlock(_FusedBlock):
    def __init__(self, in_ch, out_ch, norm_type, activation, use_se, se_reduction, down=True):
        super().__init__()
        self.stride = 2 if down else 1
        self.conv = nn.Conv2d(in_ch, out_ch, 3, stride=self.stride, padding=1)
        self.norm = _get_norm(norm_type, out_ch)
        self.act = activation
        self.se = SEBlockWrapper(out_ch, use_se, se_reduction)

    def _conv(self):
        return self.conv


class DeconvBlock(_FusedBlock):
    def __init__(self, in_ch, out_ch, norm_type, activation, use_se, se_reduction, up=True):
        super().__init__()
        self.upsample = bool(up)
        if up:
            self.up = nn.Sequential(
                nn.Upsample(scale_factor=2, mode="bilinear", align_corners=False),
                nn.Conv2d(in_ch, out_ch, kernel_size=3, stride=1, padding=1),
            )
        else:
            self.up = nn.Conv2d(in_ch, out_ch, 3, padding=1)
        self.norm = _get_norm(norm_type, out_ch)
        self.act = activation
        self.se = SEBlockWrapper(out_ch, use_se, se_reduction)

    def _conv(self):
        return self.up[1] if self.upsample else self.up


class BetaVAE(nn.Module):
    def __init__(self, config=None, precision: str = "auto"):
        super().__init__()
        self.cfg = config or get_config()
        mcfg, dcfg = self.cfg.model, self.cfg.data
        lcfg = getattr(self.cfg, "loss", None)
        self.precision = precision
        self.latent_dim = mcfg.latent_dim
        self.deterministic = getattr(mcfg, "deterministic_overfit", False)
        self.latent_reg_lambda = getattr(mcfg, "latent_reg_lambda", 0.0)
        self.pooling = getattr(mcfg, "encoder_pooling", "flatten")
        self.latent_clamp = getattr(mcfg, "latent_clamp", None)
        self.use_lpips = getattr(lcfg, "use_lpips", False)
        self.use_ffl = getattr(lcfg, "use_ffl", False)
        self.lpips_weight = getattr(lcfg, "lpips_weight", 0.0)
        self.ffl_weight = getattr(lcfg, "ffl_weight", 0.0)
        self.lpips_net_name = getattr(lcfg, "lpips_net", "alex")
        self.lpips_loss = None
        self.ffl_loss = FocalFrequencyLoss(alpha=getattr(lcfg, "ffl_alpha", 1.0))
        if self.pooling == "gap":
            raise NotImplementedError("encoder_pooling 'gap': only 'flatten' is covered by the fused heads kernel")
        if self.pooling != "flatten":
            raise ValueError(f"Unknown encoder_pooling {self.pooling}")

        act = _get_activation(mcfg.activation)
        norm_type = mcfg.encoder_norm
        in_ch = 1 if dcfg.grayscale else 3
        widths = [in_ch] + [mcfg.base_channels * (2 ** i) for i in range(mcfg.num_blocks)]
        self.encoder = nn.Sequential(*[
            ConvBlock(widths[i], widths[i + 1], norm_type, act, True, mcfg.se_reduction_ratio, down=True)
            for i in range(mcfg.num_blocks)])
        self.enc_pool = nn.Identity()
        side = dcfg.image_size
        for _ in range(mcfg.num_blocks):
            side = ops.conv_out(side, 2)
        self.enc_out_shape = torch.Size((widths[-1], side, side))  # derived arithmetically, no dummy forward
        self.flat_dim = widths[-1] * side * side
        self.fc_mu = nn.Linear(self.flat_dim, self.latent_dim)
        self.fc_logvar = nn.Linear(self.flat_dim, self.latent_dim)
        self.fc_dec = nn.Linear(self.latent_dim, self.flat_dim)
        rev = widths[:0:-1]
        self.decoder_blocks = nn.Sequential(*[
            DeconvBlock(rev[i], rev[i + 1] if i + 1 < len(rev) else rev[-1], norm_type, act,
                        mcfg.use_decoder_se, mcfg.se_reduction_ratio, up=True)
            for i in range(mcfg.num_blocks)])
        self.final_conv = nn.Conv2d(rev[-1], in_ch, 3, padding=1)
        self.current_beta = mcfg.beta
        self.recon_loss_type = mcfg.reconstruction_loss
        self.logvar_clamp = getattr(mcfg, "logvar_clamp", None)
        self._packs = ops.PackCache()
        self._eps_override = None

    # ------------------------------------------------------------------ helpers
    def set_beta(self, beta):
        self.current_beta = beta

    def set_eps(self, eps):
        """Inject the reparameterisation noise used by the next forward/loss calls (None = draw)."""
        self._eps_override = eps

    def invalidate_packs(self):
        """Call after writing parameters through ``.data`` / raw pointers."""
        self._packs.bump()

    def _dt(self):
        if self.precision == "bf16":
            return BF16
        if self.precision == "fp32":
            return F32
        return BF16 if torch.is_autocast_enabled("cuda") else F32

    def _clamp_bounds(self):
        if self.logvar_clamp:
            return float(self.logvar_clamp[0]), float(self.logvar_clamp[1])
        return -10.0, 10.0

    @staticmethod
    def _unit_f32(x):
        """fp32 image batch in [0, 1]; uint8 images (the dataset's PNG bytes) become ToTensor's floats on the device:
        x / 255, bit-exact.  Done ONCE per call, so the encoder input and the loss targets are the same tensor."""
        if not x.is_cuda:
            raise RuntimeError("BetaVAE (bvae_b200) runs on CUDA tensors only; there is no CPU path")
        return ops.u8_to_unit_f32(x.contiguous()) if x.dtype == torch.uint8 else x.contiguous().float()

    def _nhwc_input(self, x):
        x = self._unit_f32(x)
        b, c, h, w = x.shape
        return x.view(b, h, w, 1) if c == 1 else ops.nchw_to_nhwc(x, F32)

    def _conv_plan(self):
        """(cache key prefix, conv module, input H, input W, stride) for every encoder / decoder block."""
        plan = []
        side = self.cfg.data.image_size
        for i, blk in enumerate(self.encoder):
            plan.append((f"encoder.{i}", blk._conv(), side, side, 2))
            side = ops.conv_out(side, 2)
        for i, blk in enumerate(self.decoder_blocks):
            side = side * 2 if blk.upsample else side
            plan.append((f"decoder_blocks.{i}", blk._conv(), side, side, 1))
        return plan

    def _prepack(self, dt, in_dt):
        """Re-make every stale conv weight pack (fprop + dgrad layouts) with ONE kernel launch."""
        todo = []
        for idx, (key, conv, H, W, stride) in enumerate(self._conv_plan()):
            w = conv.weight
            cout, cin = w.shape[0], w.shape[1]
            first_dt = dt if (idx > 0 or in_dt == dt) else F32      # the first conv reads the fp32 image
            lf = ops.conv_weight_layout(H, W, cin, cout, stride, first_dt, 0)
            lb = ops.conv_weight_layout(H, W, cin, cout, stride, dt, 1)
            for k, transpose, layout in ((f"{key}.fprop.{lf}", False, lf), (f"{key}.dgrad.{lb}", True, lb)):
                if not self._packs.valid(k, (w,), dt):
                    todo.append((k, w, transpose, layout))
        if todo:
            with torch.no_grad():
                outs = ops.pack_conv_many([(w, tr, la) for _, w, tr, la in todo], dt)
            for (k, w, _, _), o in zip(todo, outs):
                self._packs.put(k, (w,), dt, o)

    def _encode_features(self, x, dt):
        h = self._nhwc_input(x)
        self._prepack(dt, F32)
        for i, blk in enumerate(self.encoder):
            h = blk(h, self._packs, f"encoder.{i}", dt)
        return h

    def _heads(self, h, eps, dt, beta=0.0, use_cap=False, cap=0.0, cap_w=0.0, fb=0.0, det=False):
        c, hs, ws = self.enc_out_shape
        wh = self._packs.get("heads", (self.fc_mu.weight, self.fc_logvar.weight), dt,
                             lambda: ops.pack_heads(self.fc_mu.weight, self.fc_logvar.weight, dt, c, hs * ws))
        lo, hi = self._clamp_bounds()
        return ops.HeadsReparamKL.apply(h, self.fc_mu.weight, self.fc_mu.bias, self.fc_logvar.weight,
                                        self.fc_logvar.bias, eps, wh, lo, hi, float(beta), bool(use_cap), float(cap),
                                        float(cap_w), float(fb), float(self.latent_reg_lambda), bool(det), dt)

    def _decode_features(self, z, dt):
        c, hs, ws = self.enc_out_shape
        wd, bd = self._packs.get("fc_dec", (self.fc_dec.weight, self.fc_dec.bias), dt,
                                 lambda: ops.pack_fcdec(self.fc_dec.weight, self.fc_dec.bias, dt, c, hs * ws))
        zc = float(self.latent_clamp) if self.latent_clamp is not None else 0.0
        self._prepack(dt, F32)
        h = ops.FcDec.apply(z, self.fc_dec.weight, self.fc_dec.bias, wd, bd, zc, dt, c, hs, ws)
        blocks = list(self.decoder_blocks)
        if blocks and blocks[0].upsample:
            h = ops.Upsample2x.apply(h)                     # nn.Upsample in front of the first decoder conv
        for i, blk in enumerate(blocks):
            nxt = i + 1 < len(blocks) and blocks[i + 1].upsample
            h = blk(h, self._packs, f"decoder_blocks.{i}", dt, up_after=nxt)
        return h

    def _draw_eps(self, n, device):
        if self._eps_override is not None:
            e = self._eps_override
            return e.to(device=device, dtype=torch.float32).contiguous()
        return torch.randn(n, self.latent_dim, device=device, dtype=torch.float32)

    # ------------------------------------------------------------------ reference API
    def encode(self, x):
        dt = self._dt()
        mu, logvar, _, _, _ = self._heads(self._encode_features(x, dt), None, dt, det=True)
        return mu, logvar

    def reparameterize(self, mu, logvar):
        return mu + self._draw_eps(mu.shape[0], mu.device) * torch.exp(0.5 * logvar)

    def decode(self, z):
        dt = self._dt()
        h = self._decode_features(z, dt)
        recon, _ = ops.FinalConvSigmoidRecon.apply(h, self.final_conv.weight, self.final_conv.bias, None, 0)
        return recon

    def forward(self, x, deterministic: bool = None, eps=None):
        """(recon, mu, logvar, z); ``deterministic=None`` follows ``model.deterministic_overfit``."""
        if deterministic is None:
            deterministic = self.deterministic
        dt = self._dt()
        h = self._encode_features(x, dt)
        if deterministic:
            mu, logvar, z, _, _ = self._heads(h, None, dt, det=True)
        else:
            e = eps if eps is not None else self._draw_eps(x.shape[0], x.device)
            mu, logvar, z, _, _ = self._heads(h, e.contiguous().float(), dt)
        recon, _ = ops.FinalConvSigmoidRecon.apply(self._decode_features(z, dt), self.final_conv.weight,
                                                   self.final_conv.bias, None, 0)
        return recon, mu, logvar, z

    def reconstruction_loss(self, recon, x):
        """Stand-alone form (device torch ops); ``loss()`` uses the fused kernel instead."""
        if self.recon_loss_type == "mse":
            return (recon - x).pow(2).sum() / x.size(0)
        if self.recon_loss_type == "bce":
            return torch.nn.functional.binary_cross_entropy(recon, x, reduction="sum") / x.size(0)
        if self.recon_loss_type == "l1":
            return (recon - x).abs().sum() / x.size(0)
        raise ValueError("invalid reconstruction_loss")

    def loss(self, x, beta=None, capacity=None, free_bits=0.0, capacity_weight=None, eps=None):
        if self.recon_loss_type not in LOSS:
            raise ValueError("invalid reconstruction_loss")
        dt = self._dt()
        dev = x.device
        x = self._unit_f32(x)              # encoder input, reconstruction target and FFL target are all this tensor
        b = self.current_beta if beta is None else beta
        use_capacity = (capacity is not None) and (capacity_weight is not None)
        det = bool(self.deterministic)
        e = None
        if not det:
            e = eps.to(dev).contiguous().float() if eps is not None else self._draw_eps(x.shape[0], dev)
        h = self._encode_features(x, dt)
        mu, logvar, z, kl_per_dim, sc = self._heads(
            h, e, dt, beta=float(b if b is not None else 1.0), use_cap=use_capacity,
            cap=float(capacity) if use_capacity else 0.0, cap_w=float(capacity_weight) if use_capacity else 0.0,
            fb=float(free_bits), det=det)
        feat = self._decode_features(z, dt)
        recon, base_recon = ops.FinalConvSigmoidRecon.apply(
            feat, self.final_conv.weight, self.final_conv.bias, x, LOSS[self.recon_loss_type])

        lp = ff = None
        if self.use_lpips and self.lpips_weight > 0:
            if self.lpips_loss is None:
                self.lpips_loss = LPIPSLoss(net=self.lpips_net_name).to(dev)
            lp = self.lpips_loss(recon, x) * self.lpips_weight
        if self.use_ffl and self.ffl_weight > 0:
            ff = self.ffl_loss(recon, x, weight=self.ffl_weight)
        # sc[2] = beta*kl_eff | gamma*|kl_mean - C|, plus latent_reg; the scalar tail is one launch
        total, rec_loss, packed = ops.LossCombine.apply(base_recon, lp, ff, sc)
        return {
            "total": total,
            "recon": rec_loss,
            "recon_base": base_recon.detach(),
            "recon_lpips": packed[5],
            "recon_ffl": packed[3],
            # extra keys (not in the reference's dict): the logged scalars as ONE device vector,
            # packed = [total, recon, kl_mean, recon_ffl, recon_base, recon_lpips]; scalars = its first four
            "packed": packed,
            "scalars": packed[:4],
            "kl_mean": sc[0],
            "kl_per_dim": kl_per_dim.detach(),
            "beta": torch.tensor(b),
            "capacity": torch.tensor(capacity if capacity is not None else float("nan")),
            "latent_reg": sc[3].detach() if self.latent_reg_lambda > 0 else torch.tensor(0.0),
            "recon_img": recon.detach(),
            "z": z.detach(),
            "mu": mu.detach(),
            "logvar": logvar.detach(),
            "kl_effective": sc[1].detach(),
            "mode": "capacity" if use_capacity else "beta",
        }

    def sample_prior(self, n):
        z = torch.randn(n, self.latent_dim, device=next(self.parameters()).device)
        return self.decode(z)

    def traverse(self, x, dim, steps=7, span=3.0):
        """Latent traversal; all ``steps`` decodes run as one batch (the reference loops)."""
        mu, _ = self.encode(x)
        vals = torch.linspace(-span, span, steps, device=mu.device)
        z = mu.unsqueeze(1).repeat(1, steps, 1)
        z[:, :, dim] = vals
        out = self.decode(z.view(-1, self.latent_dim))
        return out.view(mu.shape[0], steps, *out.shape[1:]), vals
