"""Auxiliary reconstruction losses -- drop-in for the reference's ``training.losses``
(reference: src/training/losses.py:5-75).

FocalFrequencyLoss is on the default config's loss path (weight 0.5).  For power-of-two image sides up to 128 (every
shipped config that switches it on) the whole term -- the 2-D FFT of the difference image in shared memory (the FFT
is linear, so F(pred) - F(target) = F(pred - target)), |F|^2, the batch-mean scale, the focal weight, the mean, and the
backward through the adjoint transform -- is ``ops.FocalFrequencyImage`` (bvae_ffl_image_fwd / _bwd): 3 + 1 launches.
Other sizes keep a library transform (``torch.fft.rfft2``, cuFFT) around the spectrum kernels
(``ops.FocalFrequencyCore``).  Its gradient reaches the fused final-conv kernel through the ``recon`` output of
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

    def forward(self, pred: torch.Tensor, target: torch.Tensor, weight: float = 1.0) -> torch.Tensor:
        """``weight`` (extra, optional): returns weight * loss from the kernel itself (saves the caller's multiply)."""
        if pred.shape != target.shape:
            raise ValueError(f"Shape mismatch: pred {pred.shape} vs target {target.shape}")
        if not pred.is_cuda:
            raise RuntimeError("FocalFrequencyLoss (bvae_b200) runs on CUDA tensors only; there is no CPU path")
        from bvae_rt import ops
        with torch.amp.autocast("cuda", enabled=False):
            if pred.dim() == 4 and ops.ffl_image_supported(pred.shape[-2], pred.shape[-1]):
                return ops.FocalFrequencyImage.apply(pred, target, self.alpha, self.eps, float(weight))
            spec = torch.fft.rfft2(pred.float() - target.float(), norm="ortho")
            loss = ops.FocalFrequencyCore.apply(torch.view_as_real(spec), pred.shape[-1], self.alpha, self.eps)
            return loss if weight == 1.0 else loss * weight


class LPIPSLoss(nn.Module):
    """Placeholder with the reference's constructor signature; LPIPS needs the third-party
    ``lpips`` package and downloaded AlexNet weights, neither available here."""

    def __init__(self, net: str = "alex", normalize: bool = True):
        super().__init__()
        raise ImportError("lpips is required for LPIPSLoss and is out of scope for the B200 hot path "
                          "(set loss.use_lpips: false).")
