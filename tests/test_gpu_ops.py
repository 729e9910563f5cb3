"""Per-kernel parity (GPU): every C-ABI op against plain fp32 PyTorch on the CPU.

fp32 kernels: rel L2 <= 1e-5.  bf16 kernels: inputs are rounded to bf16 first so the only
differences are fp32-accumulate order and the bf16 rounding of the output: rel L2 <= 1e-2."""
import pytest
import torch
import torch.nn.functional as F

from parity import rel_err

pytestmark = pytest.mark.gpu

DT = {"fp32": (torch.float32, 1e-5), "bf16": (torch.bfloat16, 1e-2)}


def _ops():
    from bvae_rt import ops
    return ops


def dev():
    return torch.device("cuda:0")


def nhwc(x, dtype):          # NCHW cpu fp32 -> NHWC cuda dtype
    return x.permute(0, 2, 3, 1).contiguous().to(dev(), dtype)


def nchw(y):                 # NHWC cuda -> NCHW cpu fp32
    return y.float().cpu().permute(0, 3, 1, 2).contiguous()


def rnd(x, dtype):           # value as the kernel sees it
    return x.to(dtype).float()


import os


class conv_path:
    """Pin the bf16 convolutions to the CUDA-core kernels ("simt") or leave the tcgen05 dispatch on ("tc")."""

    def __init__(self, which):
        self.which = which

    def __enter__(self):
        self.old = os.environ.get("BVAE_CONV_SIMT")
        os.environ["BVAE_CONV_SIMT"] = "1" if self.which == "simt" else "0"

    def __exit__(self, *a):
        if self.old is None:
            os.environ.pop("BVAE_CONV_SIMT", None)
        else:
            os.environ["BVAE_CONV_SIMT"] = self.old


CONV_SHAPES = [
    # B, H, W, Cin, Cout, stride
    (2, 8, 8, 8, 16, 1),
    (2, 8, 8, 8, 16, 2),
    (1, 7, 5, 16, 8, 1),      # odd, non-square
    (3, 9, 7, 8, 24, 2),      # odd with stride 2
    (2, 16, 16, 64, 64, 1),
    (2, 16, 16, 64, 128, 2),
    (1, 1, 1, 8, 8, 1),       # 1x1 image: only the centre tap is in range
    (2, 6, 6, 72, 40, 1),     # channel counts that are not multiples of the tile sizes
    (130, 2, 2, 8, 8, 1),     # more than one M tile from the batch dimension
    # shapes the tcgen05 kernels take in bf16 (power-of-two grids, channels 16/32/64k)
    (2, 16, 16, 64, 64, 1),
    (3, 8, 8, 128, 256, 1),   # tile spans two samples, odd batch -> zero-filled tail rows
    (2, 16, 16, 64, 128, 2),
    (1, 32, 32, 16, 32, 2),   # 32-byte swizzle (Cin = 16), 64-byte swizzle on dgrad (Cout = 32)
    (2, 4, 4, 512, 512, 1),   # 8 channel chunks per tap, two N tiles
    (1, 128, 128, 64, 64, 1), # one image row per tile
    (2, 256, 256, 16, 16, 1), # two tiles per image row
    (5, 2, 2, 64, 64, 2),     # 1x1 output grid per sample
    # stride-1 halo kernel (conv_tc_halo.cu): 16x8-pixel tiles, persistent CTAs
    (8, 128, 128, 64, 64, 1),   # resident weights, ~7 tiles per CTA (double-buffered accumulators)
    (3, 16, 8, 128, 64, 1),     # resident, two K chunks, odd tile count; dgrad: K=64 -> N=128
    (2, 32, 16, 64, 256, 1),    # resident with two N tiles per pixel tile
    (3, 16, 16, 512, 256, 1),   # streamed weights, 8 K chunks, TM=1
    (40, 32, 32, 256, 128, 1),  # streamed weights shared by TM=2 pixel tiles
    (239, 16, 8, 128, 128, 1),  # TM=2 with an odd number of tiles (last unit half empty)
    # halo wgrad on the four stride-2 parity planes / 8-row tiles
    (4, 64, 64, 64, 128, 2),
    (3, 32, 32, 128, 256, 2),
    (2, 16, 16, 256, 512, 2),
    (3, 8, 8, 64, 64, 1),
    (2, 24, 40, 64, 64, 1),
    # narrow layers (overfit_capacity.yaml): warp-level MMA wgrad (conv_wgrad_mma.cu), stride 1 and 2
    (2, 32, 16, 16, 16, 1),
    (3, 32, 32, 32, 16, 1),
    (2, 16, 16, 64, 32, 1),
    (2, 64, 32, 16, 32, 2),
    (3, 32, 32, 32, 64, 2),
    (9, 64, 64, 16, 64, 1),     # more tiles than CTAs: several tiles per CTA
    # stride-2 dgrad, halo formulation (conv_tc_s2dgrad.cu): all four parity classes per CTA / one class per CTA
    (5, 64, 32, 64, 64, 2),
    (9, 64, 64, 128, 128, 2),
    (2, 64, 64, 192, 256, 2),
]


@pytest.mark.parametrize("prec", ["fp32", "bf16", "bf16-simt"])
@pytest.mark.parametrize("shape", CONV_SHAPES)
def test_conv3x3_fprop_dgrad_wgrad(shape, prec):
    with conv_path("simt" if prec.endswith("simt") else "tc"):
        _conv_case(shape, prec.split("-")[0])


@pytest.mark.parametrize("prec", ["fp32", "bf16"])
def test_pack_conv_many_equals_single_packs(prec):
    """The one-launch tiled pack of all conv layers (bvae_pack_conv_weights: 32 x 32 x 9 tiles through shared memory, two
    elements per lane) against the element-wise single-layer pack, bit for bit: both layouts, both orientations, full
    and partial tiles, odd channel counts (single-element stores)."""
    from bvae_rt import ops
    dt = ops.BF16 if prec == "bf16" else ops.F32
    g = torch.Generator().manual_seed(3)
    shapes = [(64, 64), (128, 64), (256, 512), (48, 80), (16, 1), (33, 7), (64, 3), (5, 5)]
    jobs = []
    for cout, cin in shapes:
        w = torch.randn(cout, cin, 3, 3, generator=g).cuda()
        for tr in (False, True):
            for la in (0, 1):
                jobs.append((w, tr, la))
    many = ops.pack_conv_many(jobs, dt)
    for (w, tr, la), got in zip(jobs, many):
        assert torch.equal(got, ops.pack_conv(w, dt, tr, la)), (tuple(w.shape), tr, la)


@pytest.mark.parametrize("shape", [(2, 16, 16, 64, 64), (3, 32, 16, 128, 64), (2, 16, 32, 64, 128), (1, 32, 32, 256, 256),
                                   (33, 16, 8, 64, 64)])
def test_conv3x3_fprop_stats_epilogue(shape):
    """bvae_conv3x3_fprop_stats: the same y as the plain fprop, plus per-sample (sum, sum of squares) partials of the
    STORED (bf16) y that add up to the GroupNorm(1,C) statistics -- and GroupNorm fed with them equals GroupNorm that
    computes its own."""
    ops = _ops()
    B, H, W, Cin, Cout = shape
    g = torch.Generator().manual_seed(H + Cin + Cout)
    x = nhwc(rnd(torch.randn(B, Cin, H, W, generator=g), torch.bfloat16), torch.bfloat16)
    w = (torch.randn(Cout, Cin, 3, 3, generator=g) / (3 * Cin ** 0.5)).to(dev())
    b = torch.randn(Cout, generator=g).to(dev())
    wp = ops.pack_conv(w, ops.BF16, False, ops.conv_weight_layout(H, W, Cin, Cout, 1, ops.BF16, 0))
    y0 = ops.conv3x3_fprop(x, wp, b, Cout, 1, ops.BF16)
    y, part = ops.conv3x3_fprop(x, wp, b, Cout, 1, ops.BF16, want_stats=True)
    assert part is not None and part.dtype == torch.float64 and part.shape[0] == B and part.shape[2] == 2
    assert torch.equal(y, y0)
    yd = y.double()
    tot = part.sum(dim=1)
    assert torch.allclose(tot[:, 0], yd.sum(dim=(1, 2, 3)), rtol=1e-6, atol=1e-3)
    assert torch.allclose(tot[:, 1], yd.pow(2).sum(dim=(1, 2, 3)), rtol=1e-6, atol=1e-3)
    gamma, beta = torch.ones(Cout, device=dev()), torch.zeros(Cout, device=dev())
    import os
    old = os.environ.get("BVAE_GN_CLUSTER")
    os.environ["BVAE_GN_CLUSTER"] = "0"
    try:
        o1, s1 = ops.gn_relu_se_fwd(y, gamma, beta, None, None, None, None)
        o2, s2 = ops.gn_relu_se_fwd(y, gamma, beta, None, None, None, None, stat_part=part)
    finally:
        if old is None:
            del os.environ["BVAE_GN_CLUSTER"]
        else:
            os.environ["BVAE_GN_CLUSTER"] = old
    assert rel_err(s2[0], s1[0]) <= 1e-6 and rel_err(o2, o1) <= 1e-2 * 0.5
    # shapes the stride-1 halo kernel does not serve report "no statistics epilogue"
    from bvae_rt.lib import query
    assert query("bvae_conv3x3_fprop_stats_parts", 16, 16, 64, 64, 2, ops.BF16, ops.BF16) == 0
    assert query("bvae_conv3x3_fprop_stats_parts", 16, 16, 64, 64, 1, ops.F32, ops.F32) == 0


def _conv_case(shape, prec):
    ops = _ops()
    dtype, tol = DT[prec]
    dt = ops.BF16 if prec == "bf16" else ops.F32
    B, H, W, Cin, Cout, s = shape
    g = torch.Generator().manual_seed(B * 1000 + H * 100 + Cin + Cout + s)
    x = rnd(torch.randn(B, Cin, H, W, generator=g), dtype)
    w = torch.randn(Cout, Cin, 3, 3, generator=g) / (3 * Cin ** 0.5)
    b = torch.randn(Cout, generator=g)
    wq = rnd(w, dtype)
    xr = x.clone().requires_grad_(True)
    wr = wq.clone().requires_grad_(True)
    br = b.clone().requires_grad_(True)
    yr = F.conv2d(xr, wr, br, stride=s, padding=1)
    dy = rnd(torch.randn(yr.shape, generator=g), dtype)
    yr.backward(dy)

    wd = w.to(dev())
    wp = ops.pack_conv(wd, dt, False, ops.conv_weight_layout(H, W, Cin, Cout, s, dt, 0))
    wpt = ops.pack_conv(wd, dt, True, ops.conv_weight_layout(H, W, Cin, Cout, s, dt, 1))
    xd = nhwc(x, dtype)
    y = ops.conv3x3_fprop(xd, wp, b.to(dev()), Cout, s, dt)
    assert rel_err(nchw(y), yr) <= tol
    dyd = nhwc(dy, dtype)
    dx = ops.conv3x3_dgrad(dyd, wpt, tuple(xd.shape), s)
    assert rel_err(nchw(dx), xr.grad) <= tol
    dw, db = ops.conv3x3_wgrad(xd, dyd, s)
    assert rel_err(dw, wr.grad) <= (tol if prec == "fp32" else 1e-5 * 50)  # fp32 accumulate of bf16 inputs
    assert rel_err(db, br.grad) <= 1e-5 * (1 if prec == "fp32" else 50)


@pytest.mark.parametrize("prec", ["fp32", "bf16"])
@pytest.mark.parametrize("cin,size,cout", [(1, 16, 8), (1, 9, 16), (3, 8, 8), (1, 32, 64)])
def test_conv3x3_first_layer(cin, size, cout, prec):
    """Cin = 1 / 3 direct kernels: fp32 image in, compute-dtype activations out."""
    ops = _ops()
    dtype, tol = DT[prec]
    dt = ops.BF16 if prec == "bf16" else ops.F32
    g = torch.Generator().manual_seed(cin * 100 + size + cout)
    B = 3
    x = torch.rand(B, cin, size, size, generator=g)
    w = torch.randn(cout, cin, 3, 3, generator=g) / 3
    b = torch.randn(cout, generator=g)
    wq = rnd(w, dtype).requires_grad_(True)
    br = b.clone().requires_grad_(True)
    yr = F.conv2d(x, wq, br, stride=2, padding=1)
    dy = rnd(torch.randn(yr.shape, generator=g), dtype)
    yr.backward(dy)
    wp = ops.pack_conv(w.to(dev()), dt, False)
    xd = x.permute(0, 2, 3, 1).contiguous().to(dev())          # stays fp32
    y = ops.conv3x3_fprop(xd, wp, b.to(dev()), cout, 2, dt)
    assert y.dtype == dtype
    assert rel_err(nchw(y), yr) <= tol
    dw, db = ops.conv3x3_wgrad(xd, nhwc(dy, dtype), 2)
    assert rel_err(dw, wq.grad) <= 1e-5 * (1 if prec == "fp32" else 50)
    assert rel_err(db, br.grad) <= 1e-5 * (1 if prec == "fp32" else 50)


@pytest.mark.parametrize("prec", ["fp32", "bf16"])
@pytest.mark.parametrize("shape", [(2, 8, 8, 8), (1, 1, 1, 8), (3, 5, 7, 16), (2, 16, 16, 64), (1, 2, 1, 8)])
def test_upsample2x(shape, prec):
    ops = _ops()
    dtype, tol = DT[prec]
    B, H, W, Cc = shape
    g = torch.Generator().manual_seed(H * 10 + W)
    x = rnd(torch.randn(B, Cc, H, W, generator=g), dtype).requires_grad_(True)
    yr = F.interpolate(x, scale_factor=2, mode="bilinear", align_corners=False)
    dy = rnd(torch.randn(yr.shape, generator=g), dtype)
    yr.backward(dy)
    y = ops.upsample2x_fwd(nhwc(x.detach(), dtype))
    assert rel_err(nchw(y), yr) <= tol
    dx = ops.upsample2x_bwd(nhwc(dy, dtype))
    assert rel_err(nchw(dx), x.grad) <= tol


def _gn_se_ref(y, gamma, beta, se):
    a = torch.relu(F.group_norm(y, 1, gamma, beta, 1e-5))
    if se is None:
        return a
    w1, b1, w2, b2 = se
    s = torch.sigmoid(F.linear(torch.relu(F.linear(a.mean(dim=(2, 3)), w1, b1)), w2, b2))
    return a * s[:, :, None, None]


@pytest.mark.parametrize("path", ["stream", "cluster"])
@pytest.mark.parametrize("prec", ["fp32", "bf16"])
@pytest.mark.parametrize("shape", [(2, 8, 8, 8, 2), (3, 4, 4, 64, 8), (1, 16, 16, 16, 1), (2, 33, 9, 24, 3),
                                   (2, 64, 64, 8, 0), (4, 8, 8, 512, 64), (2, 8, 8, 16, 0),
                                   # cluster-resident kernels (gn_se_cluster.cu): 8- and 16-CTA clusters, split SE mat-vecs
                                   (2, 64, 64, 64, 8), (2, 128, 128, 64, 8), (3, 32, 32, 128, 16), (2, 16, 16, 256, 32),
                                   (2, 128, 128, 16, 0), (40, 16, 16, 64, 8),
                                   # tiled SE weight-gradient kernel: channel tail (72 = 2 * 32 + 8), batch tail (37 = 32 + 5)
                                   (3, 4, 4, 72, 40), (37, 4, 4, 128, 32)])
def test_gn_relu_se(shape, prec, path, monkeypatch):
    # "stream" = the default multi-pass kernels; "cluster" = the cluster-resident kernels (BVAE_GN_CLUSTER=1, read per call)
    monkeypatch.setenv("BVAE_GN_CLUSTER", "1" if path == "cluster" else "0")
    ops = _ops()
    dtype, tol = DT[prec]
    B, H, W, Cc, R = shape
    g = torch.Generator().manual_seed(H * 100 + Cc + R)
    y = rnd(torch.randn(B, Cc, H, W, generator=g) * 1.5 + 0.3, dtype).requires_grad_(True)
    gamma = (1 + 0.2 * torch.randn(Cc, generator=g)).requires_grad_(True)
    beta = (0.2 * torch.randn(Cc, generator=g)).requires_grad_(True)
    se = None
    if R > 0:
        se = [(torch.randn(R, Cc, generator=g) / Cc ** 0.5).requires_grad_(True),
              (0.1 * torch.randn(R, generator=g)).requires_grad_(True),
              (torch.randn(Cc, R, generator=g) / R ** 0.5).requires_grad_(True),
              (0.1 * torch.randn(Cc, generator=g)).requires_grad_(True)]
    ref = _gn_se_ref(y, gamma, beta, se)
    dout = rnd(torch.randn(ref.shape, generator=g), dtype)
    ref.backward(dout)

    D = lambda t: None if t is None else t.detach().to(dev())
    sed = [D(t) for t in se] if se else [None] * 4
    yd = nhwc(y.detach(), dtype)
    out, saved = ops.gn_relu_se_fwd(yd, D(gamma), D(beta), *sed)
    assert rel_err(nchw(out), ref) <= tol
    dy, dg, db, dw1, db1, dw2, db2 = ops.gn_relu_se_bwd(nhwc(dout, dtype), yd, D(gamma), D(beta), *sed, saved)
    ptol = 1e-5 if prec == "fp32" else 2e-3
    assert rel_err(nchw(dy), y.grad) <= tol
    assert rel_err(dg, gamma.grad) <= ptol * 3
    assert rel_err(db, beta.grad) <= ptol * 3
    if se:
        for got, want in zip((dw1, db1, dw2, db2), se):
            assert rel_err(got, want.grad) <= ptol * 5


@pytest.mark.parametrize("prec", ["fp32", "bf16"])
@pytest.mark.parametrize("shape", [(2, 8, 8, 8, 2), (3, 4, 4, 64, 8), (2, 5, 7, 24, 3), (2, 16, 16, 256, 32), (32, 8, 8, 512, 64),
                                   (3, 32, 32, 128, 16), (2, 64, 64, 64, 8), (2, 16, 16, 16, 0), (1, 1, 1, 8, 2)])
def test_gn_relu_se_fused_upsample(shape, prec, monkeypatch):
    """bvae_gn_relu_se_up_{fwd,bwd} (the decoder's activation emitted already upsampled, the transposed stencil fused into
    the backward's reduction pass) against torch -- and bit-identical to the un-fused composition of the same
    (streaming) kernels."""
    monkeypatch.setenv("BVAE_GN_CLUSTER", "0")
    ops = _ops()
    dtype, tol = DT[prec]
    B, H, W, Cc, R = shape
    g = torch.Generator().manual_seed(7 * H + Cc + R)
    y = rnd(torch.randn(B, Cc, H, W, generator=g) * 1.5 + 0.3, dtype).requires_grad_(True)
    gamma = (1 + 0.2 * torch.randn(Cc, generator=g)).requires_grad_(True)
    beta = (0.2 * torch.randn(Cc, generator=g)).requires_grad_(True)
    se = None
    if R > 0:
        se = [(torch.randn(R, Cc, generator=g) / Cc ** 0.5).requires_grad_(True),
              (0.1 * torch.randn(R, generator=g)).requires_grad_(True),
              (torch.randn(Cc, R, generator=g) / R ** 0.5).requires_grad_(True),
              (0.1 * torch.randn(Cc, generator=g)).requires_grad_(True)]
    ref = F.interpolate(_gn_se_ref(y, gamma, beta, se), scale_factor=2, mode="bilinear", align_corners=False)
    dup = rnd(torch.randn(ref.shape, generator=g), dtype)
    ref.backward(dup)
    D = lambda t: None if t is None else t.detach().to(dev())
    sed = [D(t) for t in se] if se else [None] * 4
    yd = nhwc(y.detach(), dtype)
    up, saved = ops.gn_relu_se_fwd(yd, D(gamma), D(beta), *sed, up=True)
    assert tuple(up.shape) == (B, 2 * H, 2 * W, Cc)
    # bf16: the activation is rounded to bf16 before the stencil (as the un-fused path stores it), torch keeps fp32
    assert rel_err(nchw(up), ref) <= tol * (1 if prec == "fp32" else 2)
    dupd = nhwc(dup, dtype)
    dy, dg, db, dw1, db1, dw2, db2 = ops.gn_relu_se_bwd(dupd, yd, D(gamma), D(beta), *sed, saved, up=True)
    ptol = 1e-5 if prec == "fp32" else 2e-3
    assert rel_err(nchw(dy), y.grad) <= tol * (1 if prec == "fp32" else 2)
    assert rel_err(dg, gamma.grad) <= ptol * 3
    assert rel_err(db, beta.grad) <= ptol * 3
    if se:
        for got, want in zip((dw1, db1, dw2, db2), se):
            assert rel_err(got, want.grad) <= ptol * 5
    # the fused kernels are the composition, bit for bit
    out, saved2 = ops.gn_relu_se_fwd(yd, D(gamma), D(beta), *sed)
    assert torch.equal(ops.upsample2x_fwd(out), up)
    dy2 = ops.gn_relu_se_bwd(ops.upsample2x_bwd(dupd), yd, D(gamma), D(beta), *sed, saved2)[0]
    assert torch.equal(dy2, dy)


@pytest.mark.parametrize("prec", ["fp32", "bf16"])
@pytest.mark.parametrize("cfg", [
    dict(B=3, L=8, C=16, hw=4, mode="beta", beta=1.0, fb=0.0, reg=0.0),
    dict(B=4, L=8, C=16, hw=4, mode="beta", beta=0.15, fb=0.5, reg=0.0),
    dict(B=2, L=16, C=8, hw=8, mode="cap", beta=1.0, fb=0.0, reg=0.01),
    dict(B=33, L=64, C=32, hw=8, mode="cap", beta=1.0, fb=0.0, reg=0.0),
    dict(B=2, L=8, C=16, hw=4, mode="det", beta=1.0, fb=0.0, reg=0.02),
    # batch > 128: the weight-gradient product runs several K passes (adds into the fp32 output), dh spans two row blocks
    dict(B=160, L=8, C=16, hw=4, mode="cap", beta=1.0, fb=0.0, reg=0.0),
    # F % 128 == 0 but P % 16 != 0: tensor-core path with the generic (non-patch) NHWC -> NCHW scatter
    dict(B=5, L=8, C=128, hw=3, mode="beta", beta=1.0, fb=0.0, reg=0.0),
])
def test_heads_reparam_kl(cfg, prec):
    ops = _ops()
    dtype, tol = DT[prec]
    dt = ops.BF16 if prec == "bf16" else ops.F32
    B, L, Cc, hw = cfg["B"], cfg["L"], cfg["C"], cfg["hw"]
    P, Fd = hw * hw, Cc * hw * hw
    g = torch.Generator().manual_seed(B + L + Cc)
    h = rnd(torch.randn(B, Cc, hw, hw, generator=g), dtype).requires_grad_(True)
    wmu = (torch.randn(L, Fd, generator=g) / Fd ** 0.5)
    wlv = (torch.randn(L, Fd, generator=g) * 6 / Fd ** 0.5)      # some logvars leave the clamp range
    wmu_q, wlv_q = rnd(wmu, dtype).requires_grad_(True), rnd(wlv, dtype).requires_grad_(True)
    bmu = (0.1 * torch.randn(L, generator=g)).requires_grad_(True)
    blv = (0.1 * torch.randn(L, generator=g)).requires_grad_(True)
    eps = torch.randn(B, L, generator=g)
    lo, hi, cap, capw = -2.0, 1.5, 3.0, 2.0
    det = cfg["mode"] == "det"
    flat = h.reshape(B, -1)
    mu = F.linear(flat, wmu_q, bmu)
    lv = torch.clamp(F.linear(flat, wlv_q, blv), lo, hi)
    if det:
        z, klm, kle = mu, torch.zeros(()), torch.zeros(())
        klpd = torch.zeros(L)
    else:
        z = mu + eps * torch.exp(0.5 * lv)
        kl = -0.5 * (1 + lv - mu.pow(2) - lv.exp())
        klpd = kl.mean(0)
        klm = kl.sum(1).mean()
        kle = torch.clamp(klpd, min=cfg["fb"]).sum() if (cfg["fb"] > 0 and cfg["mode"] != "cap") else klpd.sum()
    reg = cfg["reg"] * mu.pow(2).mean() if cfg["reg"] > 0 else torch.zeros(())
    if det:
        obj = reg
    elif cfg["mode"] == "cap":
        obj = capw * (klm - cap).abs() + reg
    else:
        obj = cfg["beta"] * kle + reg
    gz = torch.randn(B, L, generator=g)
    gmu = torch.randn(B, L, generator=g)
    glv = torch.randn(B, L, generator=g)
    total = obj * 0.7 + (z * gz).sum() + (mu * gmu).sum() + (lv * glv).sum() + klm * 0.3
    total.backward()

    D = lambda t: t.detach().to(dev())
    hd = nhwc(h.detach(), dtype).requires_grad_(True)
    p = [D(wmu).requires_grad_(True), D(bmu).requires_grad_(True), D(wlv).requires_grad_(True),
         D(blv).requires_grad_(True)]
    wh = ops.pack_heads(p[0].detach(), p[2].detach(), dt, Cc, P)
    out = ops.HeadsReparamKL.apply(hd, p[0], p[1], p[2], p[3], None if det else D(eps), wh, lo, hi, cfg["beta"],
                                   cfg["mode"] == "cap", cap, capw, cfg["fb"], cfg["reg"], det, dt)
    mu_d, lv_d, z_d, klpd_d, sc = out
    assert rel_err(mu_d, mu) <= tol and rel_err(lv_d, lv) <= tol and rel_err(z_d, z) <= tol
    if not det:
        assert rel_err(klpd_d, klpd) <= tol * 2
        assert rel_err(sc[0], klm) <= tol * 2 and rel_err(sc[1], kle) <= tol * 2
    assert rel_err(sc[2], obj) <= tol * 2
    tot_d = sc[2] * 0.7 + (z_d * D(gz)).sum() + (mu_d * D(gmu)).sum() + (lv_d * D(glv)).sum() + sc[0] * 0.3
    tot_d.backward()
    gt = tol if prec == "fp32" else 2e-2
    assert rel_err(nchw(hd.grad), h.grad) <= gt
    for got, want in zip(p, (wmu_q, bmu, wlv_q, blv)):
        assert rel_err(got.grad, want.grad) <= gt, rel_err(got.grad, want.grad)


@pytest.mark.parametrize("prec", ["fp32", "bf16"])
@pytest.mark.parametrize("B,L,Cc,hw,zc", [(3, 8, 16, 4, 0.0), (2, 16, 8, 8, 0.7), (33, 64, 32, 8, 0.0),
                                          (160, 8, 16, 4, 0.0), (5, 8, 128, 3, 0.5)])
def test_fc_dec(B, L, Cc, hw, zc, prec):
    ops = _ops()
    dtype, tol = DT[prec]
    dt = ops.BF16 if prec == "bf16" else ops.F32
    Fd = Cc * hw * hw
    g = torch.Generator().manual_seed(B + L)
    z = torch.randn(B, L, generator=g).requires_grad_(True)
    w = torch.randn(Fd, L, generator=g) / L ** 0.5
    wq = rnd(w, dtype).requires_grad_(True)
    b = (0.1 * torch.randn(Fd, generator=g)).requires_grad_(True)
    zz = torch.clamp(z, -zc, zc) if zc > 0 else z
    ref = F.linear(zz, wq, b).reshape(B, Cc, hw, hw)
    dh = rnd(torch.randn(ref.shape, generator=g), dtype)
    ref.backward(dh)
    D = lambda t: t.detach().to(dev())
    zd, wd_, bd_ = D(z).requires_grad_(True), D(w).requires_grad_(True), D(b).requires_grad_(True)
    wd, bd = ops.pack_fcdec(wd_.detach(), bd_.detach(), dt, Cc, hw * hw)
    h0 = ops.FcDec.apply(zd, wd_, bd_, wd, bd, zc, dt, Cc, hw, hw)
    assert rel_err(nchw(h0), ref) <= tol
    h0.backward(nhwc(dh, dtype))
    gt = tol if prec == "fp32" else 5e-3
    assert rel_err(zd.grad, z.grad) <= gt
    assert rel_err(wd_.grad, wq.grad) <= gt
    assert rel_err(bd_.grad, b.grad) <= gt


@pytest.mark.parametrize("prec", ["fp32", "bf16"])
@pytest.mark.parametrize("loss", ["mse", "l1", "bce"])
@pytest.mark.parametrize("B,H,W,Cc,Co", [(2, 8, 8, 8, 1), (3, 5, 7, 16, 1), (2, 16, 16, 64, 1), (2, 8, 8, 8, 3),
                                         (3, 32, 16, 64, 3), (5, 128, 128, 64, 1)])
def test_finalconv_sigmoid_recon(B, H, W, Cc, Co, loss, prec):
    ops = _ops()
    from bvae_rt.lib import LOSS
    dtype, tol = DT[prec]
    g = torch.Generator().manual_seed(B + H + Cc + Co)
    h = rnd(torch.randn(B, Cc, H, W, generator=g), dtype).requires_grad_(True)
    w = (torch.randn(Co, Cc, 3, 3, generator=g) / (3 * Cc ** 0.5)).requires_grad_(True)
    b = (0.1 * torch.randn(Co, generator=g)).requires_grad_(True)
    x = torch.rand(B, Co, H, W, generator=g)
    r = torch.sigmoid(F.conv2d(h, w, b, padding=1))
    if loss == "mse":
        l = F.mse_loss(r, x, reduction="sum") / B
    elif loss == "l1":
        l = F.l1_loss(r, x, reduction="sum") / B
    else:
        l = F.binary_cross_entropy(r, x, reduction="sum") / B
    gr = torch.randn(r.shape, generator=g)
    (l * 1.3 + (r * gr).sum()).backward()
    D = lambda t: t.detach().to(dev())
    hd = nhwc(h.detach(), dtype).requires_grad_(True)
    wd, bd = D(w).requires_grad_(True), D(b).requires_grad_(True)
    rd, ld = ops.FinalConvSigmoidRecon.apply(hd, wd, bd, D(x), LOSS[loss])
    # fp32: exact fp32 math.  bf16: shapes the tcgen05 forward takes (H%16 == 0, W%8 == 0, C%64 == 0) use
    # bf16-rounded weights, as the reference's conv2d does under autocast -> bf16 tolerance; the CUDA-core
    # forward keeps fp32 weights
    ftol = 1e-5 if (prec == "fp32" or H % 16 or W % 8 or Cc % 64) else 2e-3
    assert rel_err(rd, r) <= ftol and rel_err(ld, l) <= ftol
    (ld * 1.3 + (rd * D(gr)).sum()).backward()
    # l1: d|r-x| = sign(r-x) flips wherever the bf16-rounded weights move r across x (reference under autocast: same)
    assert rel_err(nchw(hd.grad), h.grad) <= (3 * tol if (loss == "l1" and ftol != 1e-5) else tol)
    gtol = 2e-5 if ftol == 1e-5 else (3 * tol if loss == "l1" else tol)   # db = sum of signed dlogit: bf16-rounded weights move it by ~1e-3..1e-2
    assert rel_err(wd.grad, w.grad) <= gtol and rel_err(bd.grad, b.grad) <= gtol
    # decode-only form (no target)
    r2, _ = ops.FinalConvSigmoidRecon.apply(hd.detach(), wd.detach(), bd.detach(), None, 0)
    assert rel_err(r2, r) <= ftol


def test_adam_clip_step_matches_torch():
    from bvae_rt import lib
    n = 100003
    g = torch.Generator().manual_seed(5)
    p0 = torch.randn(n, generator=g)
    grads = [torch.randn(n, generator=g) * (3.0 if i == 0 else 0.01) for i in range(3)]
    pr = p0.clone().requires_grad_(True)
    opt = torch.optim.Adam([pr], lr=5e-4)
    for gr in grads:
        pr.grad = gr.clone()
        torch.nn.utils.clip_grad_norm_([pr], 3.0)
        opt.step()
    d = dev()
    p, m, v = p0.to(d), torch.zeros(n, device=d), torch.zeros(n, device=d)
    ss = torch.zeros(1, device=d)
    ws = torch.empty(lib.query("bvae_adam_workspace", n), dtype=torch.uint8, device=d)
    st = torch.cuda.current_stream().cuda_stream
    for i, gr in enumerate(grads):
        gd = gr.to(d)
        lib.call("bvae_grad_sumsq", gd.data_ptr(), n, ss.data_ptr(), ws.data_ptr(), ws.numel(), st)
        assert rel_err(ss[0], gr.double().pow(2).sum().float()) <= 1e-6
        lib.call("bvae_adam_clip_step", p.data_ptr(), gd.data_ptr(), m.data_ptr(), v.data_ptr(), n, ss.data_ptr(),
                 1.0, 3.0, 5e-4, 0.9, 0.999, 1e-8, 0.0, 0, i + 1, st)
    assert rel_err(p, pr.detach()) <= 1e-6


def test_errors_are_loud():
    ops = _ops()
    from bvae_rt.lib import BvaeError
    x = torch.zeros(1, 4, 4, 6, device=dev())           # Cin = 6: neither 1/3 nor a multiple of 4
    wp = torch.zeros(9, 6, 8, device=dev())
    with pytest.raises(BvaeError):
        ops.conv3x3_fprop(x, wp, None, 8, 1, ops.F32)
    with pytest.raises(BvaeError):
        ops.upsample2x_fwd(torch.zeros(1, 2, 2, 8))       # CPU tensor: no CPU path


@pytest.mark.parametrize("alpha", [1.0, 0.5, 2.0])
@pytest.mark.parametrize("shape", [(4, 1, 32, 32), (2, 3, 16, 24), (3, 1, 8, 9), (32, 1, 128, 128),
                                   # shared-memory FFT path (power-of-two sides <= 128): non-square, multi-channel,
                                   # smallest side, and (batch 3) an image with zero difference -> clamp_min branch
                                   (2, 3, 64, 128), (3, 2, 16, 8), (3, 1, 2, 4), (5, 1, 128, 32)])
def test_focal_frequency_loss_core(shape, alpha):
    """FocalFrequencyLoss on CUDA (shared-memory FFT kernels, or rfft2 + bvae_ffl_* for other sizes) against the
    reference composition (fft2 of both images, src/training/losses.py:23-33) in fp64, values and gradient."""
    from training.losses import FocalFrequencyLoss
    g = torch.Generator().manual_seed(sum(shape))
    pred = torch.rand(shape, generator=g)
    target = torch.rand(shape, generator=g)
    if shape[0] == 3:
        pred[0] = target[0]                    # an image with zero difference: exercises the clamp_min branch

    def ref(p, t, eps=1e-8):
        spec = torch.fft.fft2(p, norm="ortho") - torch.fft.fft2(t, norm="ortho")
        power = spec.real.square() + spec.imag.square()
        scale = power.mean(dim=(0, 2, 3), keepdim=True) + eps
        return ((power / scale).clamp_min(eps).pow(alpha) * power).mean()

    pr = pred.double().requires_grad_(True)
    want = ref(pr, target.double())
    want.backward()
    pd = pred.to(dev()).requires_grad_(True)
    got = FocalFrequencyLoss(alpha=alpha)(pd, target.to(dev()))
    (got * 1.7).backward()
    assert rel_err(got, want) <= 2e-5
    assert rel_err(pd.grad / 1.7, pr.grad) <= 5e-5
    # weight folded into the kernels
    pd2 = pred.to(dev()).requires_grad_(True)
    got2 = FocalFrequencyLoss(alpha=alpha)(pd2, target.to(dev()), weight=0.5)
    got2.backward()
    assert rel_err(got2, 0.5 * want) <= 2e-5 and rel_err(pd2.grad, 0.5 * pr.grad) <= 5e-5


@pytest.mark.parametrize("B,size", [(2, 32), (3, 64), (5, 128)])
def test_conv_first_layer_tcgen05(B, size):
    """First encoder layer (1 -> 64 channels, stride 2) through the im2col + tcgen05 path against F.conv2d on the
    bf16-rounded image / weights (what conv2d sees under autocast)."""
    ops = _ops()
    if not ops.conv_first_supported(size, size, 1, 64, 2, ops.BF16):
        pytest.skip("shape not covered by the tcgen05 first-layer path")
    g = torch.Generator().manual_seed(B + size)
    x = torch.floor(torch.rand(B, 1, size, size, generator=g) * 255) / 255
    w = torch.randn(64, 1, 3, 3, generator=g) / 3
    b = torch.randn(64, generator=g) * 0.1
    xr = rnd(x, torch.bfloat16)
    wr = rnd(w, torch.bfloat16).requires_grad_(True)
    br = rnd(b, torch.bfloat16).requires_grad_(True)
    yr = F.conv2d(xr, wr, br, stride=2, padding=1)
    dy = rnd(torch.randn(yr.shape, generator=g), torch.bfloat16)
    yr.backward(dy)
    xd = x.to(dev()).view(B, size, size, 1).contiguous()
    y, xcol = ops.conv_first_fprop(xd, w.to(dev()), b.to(dev()))
    assert rel_err(nchw(y), yr) <= 1e-2
    dw, db = ops.conv_first_wgrad(xcol, nhwc(dy, torch.bfloat16))
    assert rel_err(dw, wr.grad) <= 5e-4          # fp32 accumulation of bf16 products
    assert rel_err(db, br.grad) <= 5e-4


@pytest.mark.parametrize("n", [1, 15, 16, 17, 4099, 32 * 128 * 128])
def test_u8_to_unit_f32_matches_totensor(n):
    ops = _ops()
    g = torch.Generator().manual_seed(n)
    u8 = torch.randint(0, 256, (n,), generator=g, dtype=torch.uint8)
    got = ops.u8_to_unit_f32(u8.to(dev()))
    assert torch.equal(got.cpu(), u8.float().div(255))       # bit-exact: IEEE division, as ToTensor
