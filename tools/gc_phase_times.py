#!/usr/bin/env python
"""Phase breakdown of the cluster-resident GroupNorm + ReLU + SE kernels (dev tool).

Needs a library built with -DBVAE_GC_TIMING (tools/gc_phase_times.py --build does it in place; rebuild normally
afterwards with `python __graft_entry__.py`): CTA (0, 0) of every launch records clock64() at its phase boundaries.
Prints the median cycle deltas over a few launches per layer shape of the default config (B = 32)."""
import ctypes
import glob
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
PKG = os.path.join(ROOT, "beta-vae_b200")
for p in (ROOT, os.path.join(PKG, "src"), PKG):
    sys.path.insert(0, p)

FWD = ["launch->load issued+landed", "stats pass + 2 block sums", "cluster.sync x2 + stats gather", "pool pass",
       "channel reduce", "all-gather + cluster.sync", "pool finalize", "SE hidden (W1)", "SE gate (W2)", "apply + store"]
BWD = ["load landed", "A1/A2 pass", "channel reduce", "cluster.sync x2 + gather", "A finalize", "dz1 (W2)", "dpool (W1)",
       "G1/G2 + 2 block sums", "dy pass + store"]


def build():
    import __graft_entry__ as e
    nvcc = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
    src = os.path.join(e.CSRC, "gn_se_cluster.cu")
    obj = os.path.join(PKG, "build", "gn_se_cluster.o")
    e.build()
    subprocess.check_call([nvcc, *e.NVCC_FLAGS, "-DBVAE_GC_TIMING", "-I", os.path.join(ROOT, "include"), "-c", src, "-o", obj])
    objs = sorted(glob.glob(os.path.join(PKG, "build", "*.o")))
    subprocess.check_call([nvcc, "-shared", "-o", e.LIB, *objs, "-lcudart", "-lcuda"])
    print("built", e.LIB, "with BVAE_GC_TIMING")


def main():
    if "--build" in sys.argv:
        return build()
    import numpy as np
    import torch
    from bvae_rt import lib, ops
    L = lib.load()
    L.bvae_gc_timestamps.argtypes = [ctypes.c_void_p]
    B = 32
    dev = "cuda"
    shapes = [(64, 64, 64, 8), (32, 32, 128, 16), (16, 16, 256, 32), (8, 8, 512, 64)]
    buf = (ctypes.c_longlong * 32)()
    for H, W, C, R in shapes:
        y = torch.randn(B, H, W, C, device=dev).bfloat16()
        g, b = torch.rand(C, device=dev) + 0.5, torch.randn(C, device=dev) * 0.1
        w1, b1 = torch.randn(R, C, device=dev) * 0.1, torch.zeros(R, device=dev)
        w2, b2 = torch.randn(C, R, device=dev) * 0.1, torch.zeros(C, device=dev)
        dout = torch.randn_like(y)
        f, bw = [], []
        for it in range(7):
            out, saved = ops.gn_relu_se_fwd(y, g, b, w1, b1, w2, b2)
            torch.cuda.synchronize()
            L.bvae_gc_timestamps(buf)
            t = np.array(buf[:])
            f.append(np.diff(t[0:11]))
            if False:
                print("   gate detail (TS 8, 11, 12, 13, 14, 9):", [int(t[k] - t[8]) for k in (11, 12, 13, 14, 9)])
            ops.gn_relu_se_bwd(dout, y, g, b, w1, b1, w2, b2, saved)
            torch.cuda.synchronize()
            L.bvae_gc_timestamps(buf)
            t = np.array(buf[:])
            bw.append(np.diff(t[16:26]))
        f, bw = np.median(np.array(f[2:]), 0), np.median(np.array(bw[2:]), 0)
        print(f"== {H}x{W}x{C} R={R}  (cycles of CTA (0,0); ~1.9 cycles/ns)")
        print(f"  fwd total {f.sum():7.0f} cyc = {f.sum() / 1.9e3:5.1f} us")
        for n, v in zip(FWD, f):
            print(f"     {n:34s} {v:7.0f}")
        print(f"  bwd total {bw.sum():7.0f} cyc = {bw.sum() / 1.9e3:5.1f} us")
        for n, v in zip(BWD, bw):
            print(f"     {n:34s} {v:7.0f}")


if __name__ == "__main__":
    main()
