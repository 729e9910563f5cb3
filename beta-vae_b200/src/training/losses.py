"""Auxiliary reconstruction losses -- drop-in for the reference's ``training.losses``
(reference: src/training/losses.py:5-75).

FocalFrequencyLoss is on the default config's loss path (weight 0.5) but is not one of the
subsystems BASELINE.json's north_star names; it stays a device-side torch.fft composition
("next" row 2).  Its gradient reaches the fused final-conv kernel through
the ``recon`` output of ``ops.FinalConvSigmoidRecon``.
LPIPS is a third-party AlexNet (package + weights unavailable offline): out of scope."""
import torch
import torch.nn as nn


class FocalFrequencyLoss(nn.Module):
    """mean(w * |F(pred) - F(target)|^2) with w = (|dF|^2 / mean_{n,h,w}|dF|^2)^alpha, fp32,
    orthonormal 2-D FFT; the weight is part of the graph (not detached), as in the reference."""

    def __init__(self, alpha: float = 1.0, eps: float = 1e-8):
        super().__init__()
        self.alpha = alpha
        self.eps = eps

    def forward(self, pred: torch.Tensor, target: torch.Tensor) -> torch.Tensor:
        if pred.shape != target.shape:
            raise ValueError(f"Shape mismatch: pred {pred.shape} vs target {target.shape}")
        with torch.amp.autocast("cuda", enabled=False):
            spec = torch.fft.fft2(pred.float(), norm="ortho") - torch.fft.fft2(target.float(), norm="ortho")
            power = spec.real.square() + spec.imag.square()
            scale = power.mean(dim=(0, 2, 3), keepdim=True) + self.eps
            weight = (power / scale).clamp_min(self.eps).pow(self.alpha)
            return (weight * power).mean()


class LPIPSLoss(nn.Module):
    """Placeholder with the reference's constructor signature; LPIPS needs the third-party
    ``lpips`` package and downloaded AlexNet weights, neither available here."""

    def __init__(self, net: str = "alex", normalize: bool = True):
        super().__init__()
        raise ImportError("lpips is required for LPIPSLoss and is out of scope for the B200 hot path "
                          "(set loss.use_lpips: false).")
