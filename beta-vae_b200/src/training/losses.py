"""Auxiliary reconstruction losses -- drop-in for the reference's ``training.losses``
(reference: src/training/losses.py:5-75).

FocalFrequencyLoss is on the default config's loss path (weight 0.5).  The transform is cuFFT
(``torch.fft.rfft2`` of the difference image: the FFT is linear, so F(pred) - F(target) =
F(pred - target)); everything after it -- |F|^2, the batch-mean scale, the focal weight, the mean and
their backward -- is ``ops.FocalFrequencyCore`` (bvae_ffl_fwd / bvae_ffl_bwd), four launches instead of
about twenty-five.  Its gradient reaches the fused final-conv kernel through the ``recon`` output of
``ops.FinalConvSigmoidRecon``.  There is no CPU path: CPU tensors raise.
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
        if not pred.is_cuda:
            raise RuntimeError("FocalFrequencyLoss (bvae_b200) runs on CUDA tensors only; there is no CPU path")
        from bvae_rt import ops
        with torch.amp.autocast("cuda", enabled=False):
            spec = torch.fft.rfft2(pred.float() - target.float(), norm="ortho")
            return ops.FocalFrequencyCore.apply(torch.view_as_real(spec), pred.shape[-1], self.alpha, self.eps)


class LPIPSLoss(nn.Module):
    """Placeholder with the reference's constructor signature; LPIPS needs the third-party
    ``lpips`` package and downloaded AlexNet weights, neither available here."""

    def __init__(self, net: str = "alex", normalize: bool = True):
        super().__init__()
        raise ImportError("lpips is required for LPIPSLoss and is out of scope for the B200 hot path "
                          "(set loss.use_lpips: false).")
