"""Synthetic MRI-shaped input batches, generated on the device.

Same four class patterns and value distribution as the reference's demo-data script
(reference: scripts/generate_demo_data.py:21-47): canvas clip(N(0.25, 0.08)); glioma = +0.35
inside a centred disc of radius size//4; meningioma = +0.25 on horizontal bands; pituitary =
+0.25 on a cross-hatch; notumor = extra N(0, 0.02) noise; clip to [0,1]; quantise to 1/255
(the reference writes uint8 PNGs and reads them back with ToTensor).  Classes cycle so a batch
holds them in equal proportion."""
import torch


def synthetic_batch(n: int, size: int, seed: int = 0, device="cuda", in_ch: int = 1) -> torch.Tensor:
    g = torch.Generator(device=device).manual_seed(seed)
    dev = torch.device(device)
    yy, xx = torch.meshgrid(torch.arange(size, device=dev), torch.arange(size, device=dev), indexing="ij")
    canvas = (0.25 + 0.08 * torch.randn(n, size, size, generator=g, device=dev)).clamp_(0, 1)
    extra = 0.02 * torch.randn(n, size, size, generator=g, device=dev)
    disc = ((xx - size // 2) ** 2 + (yy - size // 2) ** 2) <= (size // 4) ** 2
    bands = (yy % max(1, size // 8)) < (size // 16)
    period = max(1, size // 6)
    hatch = (((xx + yy) % period) < (size // 16)) | (((xx - yy) % period) < (size // 16))
    cls = torch.arange(n, device=dev) % 4
    add = torch.zeros(n, size, size, device=dev)
    add += (cls == 0).view(n, 1, 1) * (0.35 * disc)
    add += (cls == 1).view(n, 1, 1) * (0.25 * bands)
    add += (cls == 2).view(n, 1, 1) * (0.25 * hatch)
    add += (cls == 3).view(n, 1, 1) * extra
    # divide by a device tensor, not a Python scalar: ATen turns `x / scalar` on CUDA into x * (1 / scalar), which is
    # 1 ulp off ToTensor's (CPU, IEEE) byte / 255 for half of the 256 levels
    img = torch.floor((canvas + add).clamp_(0, 1) * 255.0) / torch.full((), 255.0, device=dev)
    return img.view(n, 1, size, size).expand(n, in_ch, size, size).contiguous()
