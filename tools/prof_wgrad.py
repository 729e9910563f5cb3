#!/usr/bin/env python
"""Dev tool: run the dec1 / dec3 weight-gradient op a few times (for an ncu capture of the wgrad + reduce kernels)."""
import os
import sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
PKG = os.path.join(ROOT, "beta-vae_b200")
for p in (ROOT, os.path.join(PKG, "src"), PKG):
    sys.path.insert(0, p)
import torch  # noqa: E402
import __graft_entry__ as entry  # noqa: E402

entry.build()
from bvae_rt import ops  # noqa: E402
dev = torch.device("cuda")
for (H, Ci, Co) in ((32, 256, 128), (128, 64, 64)):
    x = torch.randn(32, H, H, Ci, device=dev).to(torch.bfloat16)
    dy = torch.randn(32, H, H, Co, device=dev).to(torch.bfloat16)
    for _ in range(3):
        ops.conv3x3_wgrad(x, dy, 1)
torch.cuda.synchronize()
print("ok")
