#!/usr/bin/env python
"""Dev tool: where does the stride-1 halo conv kernel spend its time?  Times fprop at the decoder shapes with parts of
the kernel switched off through BVAE_HALO_DBG (1 no global stores, 2 no epilogue work, 4 no MMAs, 8 no halo TMA)."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
PKG = os.path.join(ROOT, "beta-vae_b200")
for p in (ROOT, os.path.join(PKG, "src"), PKG):
    sys.path.insert(0, p)
import torch  # noqa: E402

import __graft_entry__ as entry  # noqa: E402
from tools.bench_ops import timed_graph  # noqa: E402


def main():
    entry.build()
    from bvae_rt import ops
    from bvae_rt.lib import BF16
    B = 32
    dev = torch.device("cuda")
    shapes = [(128, 64, 64), (64, 128, 64), (32, 256, 128), (16, 512, 256)]
    for (H, Ci, Co) in shapes:
        xs = [torch.randn(B, H, H, Ci, device=dev).to(torch.bfloat16) for _ in range(3)]
        w = torch.randn(Co, Ci, 3, 3, device=dev) * 0.05
        bias = torch.zeros(Co, device=dev)
        wp = ops.pack_conv(w, BF16, False, ops.conv_weight_layout(H, H, Ci, Co, 1, BF16, 0))
        fl = 2.0 * B * H * H * Co * 9 * Ci
        mmas = B * H * H / 128 * 9 * (Ci / 16) * (Co / min(Co, 128)) / 148
        for dbg in (0, 1, 2, 4, 8, 6, 12, 14):
            os.environ["BVAE_HALO_DBG"] = str(dbg)
            i = [0]

            def f():
                i[0] += 1
                ops.conv3x3_fprop(xs[i[0] % 3], wp, bias, Co, 1, BF16)
            us = timed_graph(f, 9)
            print(f"fprop {H}x{H}x{Ci}->{Co} dbg={dbg:2d}: {us:7.1f} us  {fl / us / 1e6:7.1f} TFLOP/s  "
                  f"{us * 1965 / mmas:6.1f} cyc/MMA@1965MHz", flush=True)
        os.environ["BVAE_HALO_DBG"] = "0"


if __name__ == "__main__":
    main()
