#!/usr/bin/env python
"""Per-kernel timing of the hot-path ops at the default config's layer shapes (dev tool).

Each op is captured REPS times into one CUDA graph and replayed, so the CUDA-event time is
device time without Python / launch gaps.  Buffers are reused across repetitions (L2-warm for
tensors below the 126 MB L2) -- use ncu for DRAM traffic.
    python tools/bench_ops.py [--batch 32] [--only conv] [--reps 20]"""
import argparse
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
PKG = os.path.join(ROOT, "beta-vae_b200")
for p in (ROOT, os.path.join(PKG, "src"), PKG):
    sys.path.insert(0, p)
import torch  # noqa: E402

import __graft_entry__ as entry  # noqa: E402


def timed_graph(fn, reps):
    fn()
    torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        for _ in range(reps):
            fn()
    g.replay()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    best = 1e9
    for _ in range(3):
        e0.record()
        g.replay()
        e1.record()
        torch.cuda.synchronize()
        best = min(best, e0.elapsed_time(e1) / reps)
    return best * 1e3  # us


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--batch", type=int, default=32)
    ap.add_argument("--reps", type=int, default=20)
    ap.add_argument("--only", default="")
    args = ap.parse_args()
    entry.build()
    from bvae_rt import ops
    from bvae_rt.lib import BF16
    B = args.batch
    dev = torch.device("cuda")
    bf = torch.bfloat16
    PEAK_TF, PEAK_GB = 1404.8, 6449.4
    rows = []

    def report(name, us, flops=0.0, nbytes=0.0):
        tf = flops / us / 1e6 if flops else 0.0
        gb = nbytes / us / 1e3 if nbytes else 0.0
        rows.append((name, us, tf, gb))
        print(f"{name:44s} {us:9.1f} us  {tf:8.1f} TFLOP/s ({100 * tf / PEAK_TF:5.1f}%)  {gb:8.1f} GB/s ({100 * gb / PEAK_GB:5.1f}%)",
              flush=True)

    # (H, W, Cin, Cout, stride): spatial size is the conv INPUT
    convs = [(64, 64, 64, 128, 2), (32, 32, 128, 256, 2), (16, 16, 256, 512, 2),
             (16, 16, 512, 256, 1), (32, 32, 256, 128, 1), (64, 64, 128, 64, 1), (128, 128, 64, 64, 1)]
    if not args.only or "conv" in args.only:
        for (H, W, Ci, Co, s) in convs:
            x = torch.randn(B, H, W, Ci, device=dev).to(bf)
            w = torch.randn(Co, Ci, 3, 3, device=dev) * 0.05
            bias = torch.zeros(Co, device=dev)
            Ho = ops.conv_out(H, s)
            dy = torch.randn(B, Ho, Ho, Co, device=dev).to(bf)
            wp = ops.pack_conv(w, BF16, False, ops.conv_weight_layout(H, W, Ci, Co, s, BF16, 0))
            wpt = ops.pack_conv(w, BF16, True, ops.conv_weight_layout(H, W, Ci, Co, s, BF16, 1))
            fl = 2.0 * B * Ho * Ho * Co * 9 * Ci
            tag = f"[{H}x{W}x{Ci}->{Co},s{s}]"
            report("conv_fprop" + tag, timed_graph(lambda: ops.conv3x3_fprop(x, wp, bias, Co, s, BF16), args.reps), fl,
                   (x.numel() + dy.numel()) * 2)
            report("conv_dgrad" + tag, timed_graph(lambda: ops.conv3x3_dgrad(dy, wpt, tuple(x.shape), s), args.reps), fl,
                   (x.numel() + dy.numel()) * 2)
            report("conv_wgrad" + tag, timed_graph(lambda: ops.conv3x3_wgrad(x, dy, s), args.reps), fl,
                   (x.numel() + dy.numel()) * 2)
    gns = [(64, 64, 64), (32, 32, 128), (16, 16, 256), (8, 8, 512), (16, 16, 256), (32, 32, 128), (64, 64, 64), (128, 128, 64)]
    if not args.only or "gn" in args.only:
        for (H, W, C) in sorted(set(gns)):
            R = C // 8
            y = torch.randn(B, H, W, C, device=dev).to(bf)
            dout = torch.randn(B, H, W, C, device=dev).to(bf)
            gamma, beta = torch.ones(C, device=dev), torch.zeros(C, device=dev)
            w1, b1 = torch.randn(R, C, device=dev) * 0.1, torch.zeros(R, device=dev)
            w2, b2 = torch.randn(C, R, device=dev) * 0.1, torch.zeros(C, device=dev)
            saved = [None]

            def fwd():
                out, sv = ops.gn_relu_se_fwd(y, gamma, beta, w1, b1, w2, b2)
                saved[0] = sv
            report(f"gn_relu_se_fwd[{H}x{W}x{C}]", timed_graph(fwd, args.reps), 0, 2.0 * y.numel() * 2)
            sv = saved[0]
            report(f"gn_relu_se_bwd[{H}x{W}x{C}]",
                   timed_graph(lambda: ops.gn_relu_se_bwd(dout, y, gamma, beta, w1, b1, w2, b2, sv), args.reps), 0,
                   3.0 * y.numel() * 2)
    if not args.only or "up" in args.only:
        for (H, W, C) in [(8, 8, 512), (16, 16, 256), (32, 32, 128), (64, 64, 64)]:
            x = torch.randn(B, H, W, C, device=dev).to(bf)
            dy = torch.randn(B, 2 * H, 2 * W, C, device=dev).to(bf)
            report(f"upsample_fwd[{H}x{W}x{C}]", timed_graph(lambda: ops.upsample2x_fwd(x), args.reps), 0, 5.0 * x.numel() * 2)
            report(f"upsample_bwd[{H}x{W}x{C}]", timed_graph(lambda: ops.upsample2x_bwd(dy), args.reps), 0, 5.0 * x.numel() * 2)
    if not args.only or "final" in args.only:
        H = 128
        h = torch.randn(B, H, H, 64, device=dev).to(bf).requires_grad_(True)
        w = (torch.randn(1, 64, 3, 3, device=dev) * 0.05).requires_grad_(True)
        b = torch.zeros(1, device=dev, requires_grad=True)
        tgt = torch.rand(B, 1, H, H, device=dev)
        from bvae_rt.lib import call, query
        recon = torch.empty(B, 1, H, H, device=dev)
        loss = torch.zeros((), device=dev)
        ws = torch.empty(query("bvae_finalconv_workspace", B, H, H, 64, 1), dtype=torch.uint8, device=dev)
        st = torch.cuda.current_stream().cuda_stream
        hd, wd, bd = h.detach(), w.detach(), b.detach()
        dh = torch.empty_like(hd)
        dw, db = torch.empty_like(wd), torch.empty_like(bd)
        gl = torch.ones((), device=dev)

        def ff():
            call("bvae_finalconv_sigmoid_recon_fwd", hd.data_ptr(), wd.data_ptr(), bd.data_ptr(), tgt.data_ptr(),
                 recon.data_ptr(), loss.data_ptr(), 0, ws.data_ptr(), ws.numel(), B, H, H, 64, 1, BF16,
                 torch.cuda.current_stream().cuda_stream)

        def fb():
            call("bvae_finalconv_sigmoid_recon_bwd", hd.data_ptr(), wd.data_ptr(), recon.data_ptr(), tgt.data_ptr(), None,
                 gl.data_ptr(), dh.data_ptr(), dw.data_ptr(), db.data_ptr(), 0, ws.data_ptr(), ws.numel(), B, H, H, 64, 1,
                 BF16, torch.cuda.current_stream().cuda_stream)
        report("finalconv_sigmoid_recon_fwd", timed_graph(ff, args.reps), 2.0 * B * H * H * 576, hd.numel() * 2 + 8.0 * recon.numel())
        report("finalconv_sigmoid_recon_bwd", timed_graph(fb, args.reps), 4.0 * B * H * H * 576, 2.0 * hd.numel() * 2 + 12.0 * recon.numel())
    tot = sum(r[1] for r in rows)
    print(f"sum of listed ops: {tot / 1e3:.3f} ms")


if __name__ == "__main__":
    main()
