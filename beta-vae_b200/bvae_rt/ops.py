"""torch.autograd.Function wrappers over the C ABI (include/bvae_b200.h).

Tensors crossing these functions:
  * activations: contiguous NHWC ``[B, H, W, C]`` in the compute dtype (fp32 or bf16);
  * parameters / gradients: fp32 in the reference's own layout (what ``state_dict`` holds).
Nothing here computes on the CPU or falls back to ATen for the hot ops; every call goes to
libbvae_b200.so and raises ``BvaeError`` on failure."""
from __future__ import annotations

import torch

from . import lib, profiler
from .lib import BF16, F32, call, query


def _p(t):
    if t is None:
        return None
    if not t.is_cuda:
        raise lib.BvaeError("bvae_b200 kernels need CUDA tensors (there is no CPU path)")
    if not t.is_contiguous():
        raise lib.BvaeError("bvae_b200 kernels need contiguous tensors")
    return t.data_ptr()


def _stream():
    return torch.cuda.current_stream().cuda_stream


def _ws(nbytes, device):
    return torch.empty(max(int(nbytes), 16), dtype=torch.uint8, device=device)


def torch_dtype(dt):
    return torch.bfloat16 if dt == BF16 else torch.float32


def code_of(t):
    if t.dtype == torch.bfloat16:
        return BF16
    if t.dtype == torch.float32:
        return F32
    raise lib.BvaeError(f"unsupported dtype {t.dtype}")


def conv_out(n, stride):
    return (n - 1) // stride + 1


class GradArena:
    """Destination views for parameter gradients (slices of the engine's flat bucket buffers).
    The backward kernels write straight into them, so a bucket is complete -- and its
    all-reduce can start -- the moment its last wgrad kernel has been enqueued."""

    def __init__(self):
        self._views = {}

    def register(self, param, view):
        self._views[param.data_ptr()] = (param, view)

    def out_for(self, t):
        ent = self._views.get(t.data_ptr())
        # only hand out the slot when autograd will *assign* (grad is None); when it is going to
        # accumulate into an existing .grad the kernel must write somewhere else
        if ent is not None and ent[0].grad is None and ent[1].shape == t.shape:
            return ent[1]
        return None


_ARENA = None


def set_grad_arena(arena):
    global _ARENA
    _ARENA = arena


def grad_out(t):
    """fp32 output tensor for d(loss)/d(t), in the arena when one is installed."""
    if t is None:
        return None
    if _ARENA is not None:
        v = _ARENA.out_for(t)
        if v is not None:
            return v
    return torch.empty(t.shape, dtype=torch.float32, device=t.device)


# --------------------------------------------------------------------------------------
# weight packs (fp32 master -> compute layout), cached on the parameter version
# --------------------------------------------------------------------------------------
class PackCache:
    """Compute-layout copies of the master parameters.  An entry is rebuilt when the version
    counter or storage of any source parameter changes (optimizer steps, load_state_dict);
    ``bump()`` invalidates everything (used by the fused optimizer, which writes parameters
    through raw pointers)."""

    def __init__(self):
        self._d = {}
        self._epoch = 0

    def bump(self):
        self._epoch += 1

    def valid(self, key, params, dt):
        ent = self._d.get((key, dt))
        return ent is not None and ent[0] == (self._epoch,) + tuple((p._version, p.data_ptr()) for p in params)

    def put(self, key, params, dt, val):
        self._d[(key, dt)] = ((self._epoch,) + tuple((p._version, p.data_ptr()) for p in params), val)

    def get(self, key, params, dt, build):
        sig = (self._epoch,) + tuple((p._version, p.data_ptr()) for p in params)
        ent = self._d.get((key, dt))
        if ent is not None and ent[0] == sig:
            return ent[1]
        with torch.no_grad():
            val = build()
        self._d[(key, dt)] = (sig, val)
        return val


def conv_weight_layout(H, W, Cin, Cout, stride, dt, kind):
    """0 -> [t][k][n] pack (CUDA-core kernels), 1 -> [t][n][k] pack (tcgen05 kernels); kind 0 fprop, 1 dgrad."""
    return query("bvae_conv3x3_weight_layout", H, W, Cin, Cout, stride, dt, kind)


def pack_conv(w, dt, transpose, layout=0):
    cout, cin = w.shape[0], w.shape[1]
    k, n = (cout, cin) if transpose else (cin, cout)
    shape = (9, k, n) if layout == 0 else (9, n, k)
    wp = torch.empty(shape, dtype=torch_dtype(dt), device=w.device)
    call("bvae_pack_conv_weight", _p(w), _p(wp), cout, cin, layout, 1 if transpose else 0, dt, _stream())
    return wp


def pack_conv_many(jobs, dt):
    """jobs: list of (weight [Cout,Cin,3,3] fp32, transpose, layout).  One launch; returns the packs in order."""
    import ctypes as C
    n = len(jobs)
    outs = []
    for w, transpose, layout in jobs:
        cout, cin = w.shape[0], w.shape[1]
        k, nn = (cout, cin) if transpose else (cin, cout)
        outs.append(torch.empty((9, k, nn) if layout == 0 else (9, nn, k), dtype=torch_dtype(dt), device=w.device))
    for lo in range(0, n, 32):
        part = list(range(lo, min(n, lo + 32)))
        m = len(part)
        wa = (C.c_void_p * m)(*[_p(jobs[i][0]) for i in part])
        oa = (C.c_void_p * m)(*[_p(outs[i]) for i in part])
        co = (C.c_int * m)(*[jobs[i][0].shape[0] for i in part])
        ci = (C.c_int * m)(*[jobs[i][0].shape[1] for i in part])
        la = (C.c_int * m)(*[int(jobs[i][2]) for i in part])
        tr = (C.c_int * m)(*[1 if jobs[i][1] else 0 for i in part])
        call("bvae_pack_conv_weights", m, wa, oa, co, ci, la, tr, dt, _stream())
    return outs


def pack_heads(w_mu, w_lv, dt, C, P):
    L = w_mu.shape[0]
    wh = torch.empty((2 * L, C * P), dtype=torch_dtype(dt), device=w_mu.device)
    call("bvae_pack_heads_weight", _p(w_mu), _p(w_lv), _p(wh), L, C, P, dt, _stream())
    return wh


def pack_fcdec(w, b, dt, C, P):
    L = w.shape[1]
    wd = torch.empty((L, C * P), dtype=torch_dtype(dt), device=w.device)
    bd = torch.empty((C * P,), dtype=torch.float32, device=w.device)
    call("bvae_pack_fcdec_weight", _p(w), _p(b), _p(wd), _p(bd), L, C, P, dt, _stream())
    return wd, bd


# --------------------------------------------------------------------------------------
# raw op wrappers (no autograd) -- also what the unit tests drive
# --------------------------------------------------------------------------------------
def nchw_to_nhwc(x, dt):
    B, Cc, H, W = x.shape
    y = torch.empty((B, H, W, Cc), dtype=torch_dtype(dt), device=x.device)
    call("bvae_nchw_to_nhwc", _p(x), _p(y), B, Cc, H, W, dt, _stream())
    return y


def u8_to_unit_f32(src, out=None):
    """uint8 CUDA tensor -> fp32 / 255 (ToTensor's values, bit-exact); ``out`` may be a preallocated fp32 tensor."""
    if out is None:
        out = torch.empty(src.shape, dtype=torch.float32, device=src.device)
    call("bvae_u8_to_unit_f32", _p(src), _p(out), src.numel(), _stream())
    return out


def conv3x3_fprop(x, wp, bias, cout, stride, dt, want_stats=False):
    """``want_stats``: also return the GroupNorm statistic partials [B, n, 2] float64 the kernel's epilogue produces
    (None when the kernel serving this shape has no such epilogue)."""
    B, H, W, Cin = x.shape
    y = torch.empty((B, conv_out(H, stride), conv_out(W, stride), cout), dtype=torch_dtype(dt), device=x.device)
    nparts = query("bvae_conv3x3_fprop_stats_parts", H, W, Cin, cout, stride, code_of(x), dt) if want_stats else 0
    with profiler.region(f"conv_fprop[{H}x{W}x{Cin}->{cout},s{stride}]", flops=2.0 * y.numel() * 9 * Cin,
                         nbytes=x.numel() * x.element_size() + y.numel() * y.element_size()):
        if nparts > 0:
            part = torch.empty((B, nparts, 2), dtype=torch.float64, device=x.device)
            call("bvae_conv3x3_fprop_stats", _p(x), _p(wp), _p(bias), _p(y), _p(part), B, H, W, Cin, cout, stride,
                 code_of(x), dt, _stream())
        else:
            part = None
            call("bvae_conv3x3_fprop", _p(x), _p(wp), _p(bias), _p(y), B, H, W, Cin, cout, stride, code_of(x), dt,
                 _stream())
    return (y, part) if want_stats else y


def conv3x3_dgrad(dy, wpt, in_shape, stride):
    B, H, W, Cin = in_shape
    Cout = dy.shape[3]
    dx = torch.empty(in_shape, dtype=dy.dtype, device=dy.device)
    with profiler.region(f"conv_dgrad[{H}x{W}x{Cin}<-{Cout},s{stride}]", flops=2.0 * dy.numel() * 9 * Cin,
                         nbytes=(dx.numel() + dy.numel()) * dy.element_size()):
        call("bvae_conv3x3_dgrad", _p(dy), _p(wpt), _p(dx), B, H, W, Cin, Cout, stride, code_of(dy), _stream())
    return dx


def conv3x3_wgrad(x, dy, stride, want_bias=True, dw=None, db=None):
    B, H, W, Cin = x.shape
    Cout = dy.shape[3]
    if dw is None:
        dw = torch.empty((Cout, Cin, 3, 3), dtype=torch.float32, device=x.device)
    if db is None and want_bias:
        db = torch.empty((Cout,), dtype=torch.float32, device=x.device)
    n = query("bvae_conv3x3_wgrad_workspace", B, H, W, Cin, Cout, stride)
    ws = _ws(n, x.device)
    with profiler.region(f"conv_wgrad[{H}x{W}x{Cin}->{Cout},s{stride}]", flops=2.0 * dy.numel() * 9 * Cin,
                         nbytes=x.numel() * x.element_size() + dy.numel() * dy.element_size()):
        call("bvae_conv3x3_wgrad", _p(x), _p(dy), _p(dw), _p(db), _p(ws), ws.numel(), B, H, W, Cin, Cout, stride,
             code_of(x), code_of(dy), _stream())
    return dw, db


def conv_first_supported(H, W, Cin, Cout, stride, dt):
    return dt == BF16 and query("bvae_conv_first_supported", H, W, Cin, Cout, stride) == 1


def conv_first_fprop(x, w, bias):
    """First encoder layer on tcgen05: x [B,H,W,1] fp32 -> (y [B,H/2,W/2,64] bf16, xcol) where xcol is the 16-column
    bf16 im2col (9 taps + ones column) that the weight gradient reuses."""
    B, H, W, _ = x.shape
    Ho, Wo = H // 2, W // 2
    xcol = torch.empty((B, Ho, Wo, 16), dtype=torch.bfloat16, device=x.device)
    y = torch.empty((B, Ho, Wo, 64), dtype=torch.bfloat16, device=x.device)
    ws = _ws(query("bvae_conv_first_workspace"), x.device)
    with profiler.region(f"conv_fprop[{H}x{W}x1->64,s2]", flops=2.0 * y.numel() * 9, nbytes=x.numel() * 4 + y.numel() * 2):
        call("bvae_conv_first_im2col", _p(x), _p(xcol), B, H, W, _stream())
        call("bvae_conv_first_fprop", _p(xcol), _p(w), _p(bias), _p(y), _p(ws), ws.numel(), B, Ho, Wo, _stream())
    return y, xcol


def conv_first_wgrad(xcol, dy, dw=None, db=None):
    B, Ho, Wo, _ = xcol.shape
    if dw is None:
        dw = torch.empty((64, 1, 3, 3), dtype=torch.float32, device=dy.device)
    if db is None:
        db = torch.empty((64,), dtype=torch.float32, device=dy.device)
    ws = _ws(query("bvae_conv_first_workspace"), dy.device)
    with profiler.region(f"conv_wgrad[{2 * Ho}x{2 * Wo}x1->64,s2]", flops=2.0 * dy.numel() * 9,
                         nbytes=xcol.numel() * 2 + dy.numel() * 2):
        call("bvae_conv_first_wgrad", _p(xcol), _p(dy), _p(dw), _p(db), _p(ws), ws.numel(), B, Ho, Wo, _stream())
    return dw, db


def upsample2x_fwd(x):
    B, H, W, Cc = x.shape
    y = torch.empty((B, 2 * H, 2 * W, Cc), dtype=x.dtype, device=x.device)
    with profiler.region(f"upsample_fwd[{H}x{W}x{Cc}]", nbytes=(x.numel() + y.numel()) * x.element_size()):
        call("bvae_upsample2x_fwd", _p(x), _p(y), B, H, W, Cc, code_of(x), _stream())
    return y


def upsample2x_bwd(dy):
    B, H2, W2, Cc = dy.shape
    dx = torch.empty((B, H2 // 2, W2 // 2, Cc), dtype=dy.dtype, device=dy.device)
    with profiler.region(f"upsample_bwd[{H2 // 2}x{W2 // 2}x{Cc}]", nbytes=(dx.numel() + dy.numel()) * dy.element_size()):
        call("bvae_upsample2x_bwd", _p(dy), _p(dx), B, H2 // 2, W2 // 2, Cc, code_of(dy), _stream())
    return dx


def gn_relu_se_fwd(y, gamma, beta, w1, b1, w2, b2, up=False, stat_part=None):
    """GroupNorm(1,C) + ReLU + SE.  ``up``: return the bilinear x2 upsample of the activation ([B,2H,2W,C]) without
    materialising it at its own resolution.  ``stat_part``: [B, n, 2] float64 partial (sum, sum of squares) of ``y``
    from the convolution's epilogue (skips the statistics pass)."""
    B, H, W, Cc = y.shape
    P = H * W
    dev = y.device
    has_se = w1 is not None
    R = w1.shape[0] if has_se else 0
    out = torch.empty((B, 2 * H, 2 * W, Cc), dtype=y.dtype, device=dev) if up else torch.empty_like(y)
    stats = torch.empty((B, 2), dtype=torch.float32, device=dev)
    f = lambda *s: torch.empty(s, dtype=torch.float32, device=dev)
    cpos, spos, hidden, gate = (f(B, Cc), f(B, Cc), f(B, R), f(B, Cc)) if has_se else (None,) * 4
    nsp = 0 if stat_part is None else stat_part.shape[1]
    # algorithmic bytes: one read of y + one write of out
    if up:
        ws = _ws(query("bvae_gn_relu_se_up_workspace", B, H, W, Cc), dev)
        with profiler.region(f"gn_relu_se_up_fwd[{H}x{W}x{Cc}]", nbytes=(y.numel() + out.numel()) * y.element_size()):
            call("bvae_gn_relu_se_up_fwd", _p(y), _p(out), _p(gamma), _p(beta), _p(w1), _p(b1), _p(w2), _p(b2), B, H, W,
                 Cc, R, _p(stats), _p(cpos), _p(spos), _p(hidden), _p(gate), _p(stat_part), nsp, _p(ws), ws.numel(),
                 code_of(y), _stream())
    else:
        ws = _ws(query("bvae_gn_relu_se_workspace", B, P, Cc), dev)
        with profiler.region(f"gn_relu_se_fwd[{H}x{W}x{Cc}]", nbytes=2.0 * y.numel() * y.element_size()):
            call("bvae_gn_relu_se_fwd", _p(y), _p(out), _p(gamma), _p(beta), _p(w1), _p(b1), _p(w2), _p(b2), B, P, Cc, R,
                 _p(stats), _p(cpos), _p(spos), _p(hidden), _p(gate), _p(stat_part), nsp, _p(ws), ws.numel(),
                 code_of(y), _stream())
    return out, (stats, cpos, spos, hidden, gate)


# Parameter-gradient reductions of the GroupNorm / SE blocks, deferred to the end of the backward pass: per block they
# are a ~5 us kernel that nothing but the optimizer waits for, but launched where they arise they sit on the critical
# path eight times.  ConvGnSeBlock.backward queues one job per block; an autograd final callback (runs when the whole
# backward has been enqueued, inside a CUDA-graph capture too) launches them together.  The queued entries hold the
# workspaces alive until then.  The engine switches this off when it exchanges gradient buckets from hooks DURING
# backward (the reduction would then run after its bucket had left).
_GN_DEFER = {"enabled": True, "jobs": [], "keep": []}


def set_gn_param_deferral(enabled: bool):
    _GN_DEFER["enabled"] = bool(enabled)


def flush_gn_param_jobs():
    jobs, _GN_DEFER["jobs"] = _GN_DEFER["jobs"], []
    keep, _GN_DEFER["keep"] = _GN_DEFER["keep"], []
    if jobs:
        blob = b"".join(jobs)
        with profiler.region("gn_bwd_params_batched"):
            call("bvae_gn_bwd_params_batched", blob, len(jobs), _stream())
    del keep


def gn_relu_se_bwd(dout, y, gamma, beta, w1, b1, w2, b2, saved, up=False, defer=False):
    """``up``: ``dout`` is the gradient w.r.t. the upsampled activation, [B,2H,2W,C].  ``defer`` (only inside an autograd
    backward): the parameter gradients are filled by one batched launch when the backward pass ends."""
    B, H, W, Cc = y.shape
    P = H * W
    dev = y.device
    stats, cpos, spos, hidden, gate = saved
    has_se = w1 is not None
    R = w1.shape[0] if has_se else 0
    dy = torch.empty_like(y)
    dgamma, dbeta = grad_out(gamma), grad_out(beta)
    dw1, db1, dw2, db2 = (grad_out(w1), grad_out(b1), grad_out(w2), grad_out(b2)) if has_se else (None,) * 4
    if up:
        assert tuple(dout.shape) == (B, 2 * H, 2 * W, Cc), (dout.shape, y.shape)
        ws = _ws(query("bvae_gn_relu_se_up_workspace", B, H, W, Cc), dev)
        name, nbytes = f"gn_relu_se_up_bwd[{H}x{W}x{Cc}]", (dout.numel() + 2.0 * y.numel()) * y.element_size()
    else:
        ws = _ws(query("bvae_gn_relu_se_workspace", B, P, Cc), dev)
        # algorithmic bytes: read dout + read y + write dy
        name, nbytes = f"gn_relu_se_bwd[{H}x{W}x{Cc}]", 3.0 * y.numel() * y.element_size()
    # deferral needs autograd to ADOPT the returned gradient tensors (it then sees the late writes): true when it assigns
    # (.grad is None) outside double-backward mode; when it would accumulate or clone, reduce now
    if (defer and _GN_DEFER["enabled"] and not torch.is_grad_enabled()
            and all(t is None or t.grad is None for t in (gamma, beta, w1, b1, w2, b2))):
        import ctypes
        job = ctypes.create_string_buffer(query("bvae_gn_params_job_bytes"))
        with profiler.region(name, nbytes=nbytes):
            call("bvae_gn_relu_se_bwd_deferred", _p(dout), _p(y), _p(dy), _p(gamma), _p(beta), _p(w1), _p(w2), B, H, W, Cc,
                 R, 1 if up else 0, _p(stats), _p(cpos), _p(spos), _p(hidden), _p(gate), _p(dgamma), _p(dbeta), _p(dw1),
                 _p(db1), _p(dw2), _p(db2), _p(ws), ws.numel(), code_of(y), job, _stream())
        first = not _GN_DEFER["jobs"]
        _GN_DEFER["jobs"].append(job.raw)
        # (not the gradient tensors: a second reference makes AccumulateGrad clone them -- still unfilled -- instead of
        #  adopting them; autograd / param.grad keep them alive)
        _GN_DEFER["keep"].append((ws, gamma, beta, stats, cpos, spos, hidden))
        if first:
            torch.autograd.Variable._execution_engine.queue_callback(flush_gn_param_jobs)
        return dy, dgamma, dbeta, dw1, db1, dw2, db2
    with profiler.region(name, nbytes=nbytes):
        if up:
            call("bvae_gn_relu_se_up_bwd", _p(dout), _p(y), _p(dy), _p(gamma), _p(beta), _p(w1), _p(b1), _p(w2), _p(b2), B,
                 H, W, Cc, R, _p(stats), _p(cpos), _p(spos), _p(hidden), _p(gate), _p(dgamma), _p(dbeta), _p(dw1),
                 _p(db1), _p(dw2), _p(db2), _p(ws), ws.numel(), code_of(y), _stream())
        else:
            call("bvae_gn_relu_se_bwd", _p(dout), _p(y), _p(dy), _p(gamma), _p(beta), _p(w1), _p(b1), _p(w2), _p(b2), B,
                 P, Cc, R, _p(stats), _p(cpos), _p(spos), _p(hidden), _p(gate), _p(dgamma), _p(dbeta), _p(dw1), _p(db1),
                 _p(dw2), _p(db2), _p(ws), ws.numel(), code_of(y), _stream())
    return dy, dgamma, dbeta, dw1, db1, dw2, db2


# --------------------------------------------------------------------------------------
# autograd Functions
# --------------------------------------------------------------------------------------
class Upsample2x(torch.autograd.Function):
    """nn.Upsample(x2, bilinear, align_corners=False) on NHWC (reference: src/models/beta_vae.py:52-53) -- the stand-alone
    form, used in front of the first decoder block; between decoder blocks the upsample is fused into ConvGnSeBlock."""

    @staticmethod
    def forward(ctx, x):
        return upsample2x_fwd(x)

    @staticmethod
    def backward(ctx, dy):
        return upsample2x_bwd(dy.contiguous())


class ConvGnSeBlock(torch.autograd.Function):
    """conv3x3 -> GroupNorm(1,C) -> ReLU -> [SE] [-> bilinear x2]: one encoder ConvBlock or decoder DeconvBlock of the
    reference (src/models/beta_vae.py:33-66).  A decoder block's nn.Upsample belongs to its INPUT in the reference;
    here it is produced by the block in front (``up_after``), fused into that block's normalise / rescale pass, so the
    activation is written once, already upsampled, and the backward's transposed stencil is fused into the
    GroupNorm-backward reduction pass."""

    @staticmethod
    def forward(ctx, x, cw, cb, gw, gb, w1, b1, w2, b2, packs, key, stride, up_after, dt):
        _, H, W, Cin = x.shape
        Cout = cw.shape[0]
        xin = x
        first_tc = (xin.dtype == torch.float32 and not x.requires_grad and
                    conv_first_supported(H, W, Cin, Cout, stride, dt))
        if first_tc:
            # first encoder layer: 9-tap im2col of the image + one tcgen05 MMA per tile; xcol replaces x for the backward
            y, xcol = conv_first_fprop(xin, cw, cb)
            wpt = None
            xin = xcol
        else:
            lf = conv_weight_layout(H, W, Cin, Cout, stride, dt if code_of(xin) == dt else F32, 0)
            lb = conv_weight_layout(H, W, Cin, Cout, stride, dt, 1)
            wp = packs.get(f"{key}.fprop.{lf}", (cw,), dt, lambda: pack_conv(cw, dt, False, lf))
            wpt = packs.get(f"{key}.dgrad.{lb}", (cw,), dt, lambda: pack_conv(cw, dt, True, lb))
            Ho, Wo = conv_out(H, stride), conv_out(W, stride)
            streaming = not query("bvae_gn_relu_se_prefers_unfused", x.shape[0], Ho * Wo, Cout,
                                  w1.shape[0] if w1 is not None else 0, dt)
            # where the streaming GroupNorm passes run, the conv's epilogue delivers their statistics
            y, stat_part = conv3x3_fprop(xin, wp, cb, cw.shape[0], stride, dt, want_stats=True) if streaming else \
                (conv3x3_fprop(xin, wp, cb, cw.shape[0], stride, dt), None)
        if first_tc:
            stat_part = None
        # fuse the next block's bilinear x2 only where the streaming GroupNorm passes run anyway; shapes served by the
        # single-launch cluster kernel keep the upsample as its own pass (measured faster, see gn_se.cu)
        Bn, Ho, Wo, _ = y.shape
        fuse_up = bool(up_after) and not query("bvae_gn_relu_se_prefers_unfused", Bn, Ho * Wo, Cout,
                                                 w1.shape[0] if w1 is not None else 0, code_of(y))
        out, saved = gn_relu_se_fwd(y, gw, gb, w1, b1, w2, b2, up=fuse_up, stat_part=stat_part)
        if up_after and not fuse_up:
            out = upsample2x_fwd(out)
        stats, cpos, spos, hidden, gate = saved
        ctx.save_for_backward(xin, y, gw, gb, w1, b1, w2, b2, wpt, stats, cpos, spos, hidden, gate, cw, cb)
        ctx.stride, ctx.up_after, ctx.fuse_up, ctx.first_tc = stride, up_after, fuse_up, first_tc
        return out

    @staticmethod
    def backward(ctx, dout):
        xin, y, gw, gb, w1, b1, w2, b2, wpt, stats, cpos, spos, hidden, gate, cw, cb = ctx.saved_tensors
        dout = dout.contiguous()
        if ctx.up_after and not ctx.fuse_up:
            dout = upsample2x_bwd(dout)
        dy, dgw, dgb, dw1, db1, dw2, db2 = gn_relu_se_bwd(dout, y, gw, gb, w1, b1, w2, b2,
                                                          (stats, cpos, spos, hidden, gate), up=ctx.fuse_up, defer=True)
        if ctx.first_tc:
            if ctx.needs_input_grad[0]:
                raise lib.BvaeError("the tcgen05 first-layer path does not produce an input gradient")
            dcw, dcb = conv_first_wgrad(xin, dy, dw=grad_out(cw), db=grad_out(cb))
            return None, dcw, dcb, dgw, dgb, dw1, db1, dw2, db2, None, None, None, None, None
        dcw, dcb = conv3x3_wgrad(xin, dy, ctx.stride, dw=grad_out(cw), db=grad_out(cb))
        dx = None
        if ctx.needs_input_grad[0]:
            dx = conv3x3_dgrad(dy, wpt, tuple(xin.shape), ctx.stride)
        return dx, dcw, dcb, dgw, dgb, dw1, db1, dw2, db2, None, None, None, None, None


class HeadsReparamKL(torch.autograd.Function):
    """fc_mu/fc_logvar + clamp + z = mu + eps*exp(logvar/2) + KL statistics + KL-side objective
    (src/models/beta_vae.py:137-150, 219-253).  Returns (mu, logvar, z, kl_per_dim, scalars)
    with scalars = [kl_mean, kl_effective, objective, latent_reg]."""

    @staticmethod
    def forward(ctx, h, w_mu, b_mu, w_lv, b_lv, eps, wh, lo, hi, beta, use_cap, cap, cap_w, fb, reg, det, dt):
        B, Hs, Ws, Cc = h.shape
        P, F = Hs * Ws, Hs * Ws * Cc
        L = w_mu.shape[0]
        dev = h.device
        f = lambda *s: torch.empty(s, dtype=torch.float32, device=dev)
        mu, logvar, z, klpd, sc = f(B, L), f(B, L), f(B, L), f(L), f(4)
        lv_pass = torch.empty((B, L), dtype=torch.uint8, device=dev)
        ws = _ws(query("bvae_heads_workspace", B, L, F), dev)
        with profiler.region("heads_reparam_kl_fwd", flops=4.0 * B * L * F,
                             nbytes=(h.numel() + wh.numel()) * h.element_size()):
            call("bvae_heads_reparam_kl_fwd", _p(h), _p(wh), _p(b_mu), _p(b_lv), _p(eps), lo, hi, beta, int(use_cap),
                 cap, cap_w, fb, reg, int(det), _p(mu), _p(logvar), _p(z), _p(klpd), _p(sc), _p(lv_pass), _p(ws),
                 ws.numel(), B, L, F, dt, _stream())
        ctx.save_for_backward(h, wh, eps, mu, logvar, klpd, sc, lv_pass, w_mu, b_mu, w_lv, b_lv)
        ctx.cfg = (beta, int(use_cap), cap, cap_w, fb, reg, int(det), dt, Cc, P, L)
        ctx.mark_non_differentiable(klpd)
        ctx.set_materialize_grads(False)
        return mu, logvar, z, klpd, sc

    @staticmethod
    def backward(ctx, g_mu, g_lv, g_z, _g_klpd, g_sc):
        h, wh, eps, mu, logvar, klpd, sc, lv_pass, w_mu, b_mu, w_lv, b_lv = ctx.saved_tensors
        beta, use_cap, cap, cap_w, fb, reg, det, dt, Cc, P, L = ctx.cfg
        B = h.shape[0]
        dev = h.device
        c = lambda t: None if t is None else t.contiguous().float()
        g_mu, g_lv, g_z, g_sc = c(g_mu), c(g_lv), c(g_z), c(g_sc)
        dh = torch.empty_like(h) if ctx.needs_input_grad[0] else None
        dwm, dwl, dbm, dbl = grad_out(w_mu), grad_out(w_lv), grad_out(b_mu), grad_out(b_lv)
        ws = _ws(query("bvae_heads_workspace", B, L, Cc * P), dev)
        with profiler.region("heads_reparam_kl_bwd", flops=8.0 * B * L * Cc * P,
                             nbytes=(2 * h.numel() + wh.numel()) * h.element_size() + 8.0 * L * Cc * P):
            call("bvae_heads_reparam_kl_bwd", _p(h), _p(wh), _p(eps), _p(mu), _p(logvar), _p(klpd), _p(sc),
                 _p(lv_pass), _p(g_mu), _p(g_lv), _p(g_z), _p(g_sc), beta, use_cap, cap, cap_w, fb, reg, det, _p(dh),
                 _p(dwm), _p(dwl), _p(dbm), _p(dbl), _p(ws), ws.numel(), B, L, Cc, P, dt, _stream())
        return (dh, dwm, dbm, dwl, dbl) + (None,) * 12


class FcDec(torch.autograd.Function):
    """Optional latent clamp + fc_dec, output NHWC [B, Hs, Ws, C] (src/models/beta_vae.py:153-161)."""

    @staticmethod
    def forward(ctx, z, w, b, wd, bd, z_clamp, dt, Cc, Hs, Ws):
        B, L = z.shape
        z = z.contiguous().float()
        h0 = torch.empty((B, Hs, Ws, Cc), dtype=torch_dtype(dt), device=z.device)
        with profiler.region("fc_dec_fwd", flops=2.0 * B * L * Cc * Hs * Ws,
                             nbytes=(h0.numel() + wd.numel()) * h0.element_size()):
            call("bvae_fc_dec_fwd", _p(z), _p(wd), _p(bd), _p(h0), z_clamp, B, L, Cc * Hs * Ws, dt, _stream())
        ctx.save_for_backward(z, wd, w, b)
        ctx.cfg = (z_clamp, dt, Cc, Hs * Ws)
        return h0

    @staticmethod
    def backward(ctx, dh0):
        z, wd, w, b = ctx.saved_tensors
        z_clamp, dt, Cc, P = ctx.cfg
        B, L = z.shape
        dev = z.device
        dh0 = dh0.contiguous()
        dz = torch.empty((B, L), dtype=torch.float32, device=dev)
        dw, db = grad_out(w), grad_out(b)
        ws = _ws(query("bvae_fc_dec_workspace", B, L, Cc * P), dev)
        with profiler.region("fc_dec_bwd", flops=4.0 * B * L * Cc * P,
                             nbytes=(dh0.numel() + wd.numel()) * dh0.element_size() + 4.0 * L * Cc * P):
            call("bvae_fc_dec_bwd", _p(dh0), _p(z), _p(wd), z_clamp, _p(dz), _p(dw), _p(db), _p(ws), ws.numel(), B, L,
                 Cc, P, dt, _stream())
        return (dz, dw, db) + (None,) * 7


class FinalConvSigmoidRecon(torch.autograd.Function):
    """final_conv + sigmoid + reconstruction loss (src/models/beta_vae.py:163-164, 185-192).
    Returns (recon [B,Co,H,W] fp32, loss scalar = sum(elementwise)/B)."""

    @staticmethod
    def forward(ctx, h, w, b, target, loss_type):
        B, H, W, Cc = h.shape
        Co = w.shape[0]
        dev = h.device
        recon = torch.empty((B, Co, H, W), dtype=torch.float32, device=dev)
        # the loss scalar is written by the kernel whenever a target is given (no fill launch on the training path)
        loss = torch.empty((), dtype=torch.float32, device=dev) if target is not None else \
            torch.zeros((), dtype=torch.float32, device=dev)
        if target is not None:
            target = target.contiguous().float()
        ws = _ws(query("bvae_finalconv_workspace", B, H, W, Cc, Co), dev)
        # algorithmic bytes: read h once, read target, write recon
        with profiler.region("finalconv_sigmoid_recon_fwd", flops=2.0 * recon.numel() * 9 * Cc,
                             nbytes=h.numel() * h.element_size() + 8.0 * recon.numel()):
            call("bvae_finalconv_sigmoid_recon_fwd", _p(h), _p(w), _p(b), _p(target), _p(recon), _p(loss), loss_type,
                 _p(ws), ws.numel(), B, H, W, Cc, Co, code_of(h), _stream())
        ctx.save_for_backward(h, w, recon, target, b)
        ctx.loss_type = loss_type
        ctx.set_materialize_grads(False)
        if target is None:
            ctx.mark_non_differentiable(loss)
        return recon, loss

    @staticmethod
    def backward(ctx, g_recon, g_loss):
        h, w, recon, target, b = ctx.saved_tensors
        B, H, W, Cc = h.shape
        Co = w.shape[0]
        dev = h.device
        if g_recon is not None:
            g_recon = g_recon.contiguous().float()
        if g_loss is not None:
            g_loss = g_loss.contiguous().float()
        dh = torch.empty_like(h) if ctx.needs_input_grad[0] else None
        dw, db = grad_out(w), grad_out(b)
        ws = _ws(query("bvae_finalconv_workspace", B, H, W, Cc, Co), dev)
        # algorithmic bytes: read h, write dh, read recon/target/g_recon
        with profiler.region("finalconv_sigmoid_recon_bwd", flops=4.0 * recon.numel() * 9 * Cc,
                             nbytes=2.0 * h.numel() * h.element_size() + 12.0 * recon.numel()):
            call("bvae_finalconv_sigmoid_recon_bwd", _p(h), _p(w), _p(recon), _p(target), _p(g_recon), _p(g_loss),
                 _p(dh), _p(dw), _p(db), ctx.loss_type, _p(ws), ws.numel(), B, H, W, Cc, Co, code_of(h), _stream())
        return dh, dw, db, None, None


class FocalFrequencyCore(torch.autograd.Function):
    """Spectrum-side arithmetic of the focal-frequency loss (src/training/losses.py:23-33) on the one-sided rfft2
    spectrum, passed as its real view [B, C, H, W/2+1, 2]."""

    @staticmethod
    def forward(ctx, spec_ri, W, alpha, eps):
        B, Cc, H, Wh, two = spec_ri.shape
        assert two == 2 and Wh == W // 2 + 1
        spec_ri = spec_ri.contiguous()
        dev = spec_ri.device
        loss = torch.empty((), dtype=torch.float32, device=dev)
        scale = torch.empty(Cc, dtype=torch.float32, device=dev)
        U = torch.empty(Cc, dtype=torch.float32, device=dev)
        ws = _ws(query("bvae_ffl_workspace", Cc), dev)
        with profiler.region("ffl_fwd", nbytes=2.0 * spec_ri.numel() * 4):
            call("bvae_ffl_fwd", _p(spec_ri), B, Cc, H, Wh, W, float(alpha), float(eps), _p(loss), _p(scale), _p(U),
                 _p(ws), ws.numel(), _stream())
        ctx.save_for_backward(spec_ri, scale, U)
        ctx.cfg = (W, float(alpha), float(eps))
        return loss

    @staticmethod
    def backward(ctx, g_loss):
        spec_ri, scale, U = ctx.saved_tensors
        W, alpha, eps = ctx.cfg
        B, Cc, H, Wh, _ = spec_ri.shape
        g = torch.empty_like(spec_ri)
        g_loss = g_loss.contiguous().float()
        with profiler.region("ffl_bwd", nbytes=2.0 * spec_ri.numel() * 4):
            call("bvae_ffl_bwd", _p(spec_ri), _p(scale), _p(U), _p(g_loss), _p(g), B, Cc, H, Wh, W, alpha, eps, _stream())
        return g, None, None, None


class FocalFrequencyImage(torch.autograd.Function):
    """weight * FocalFrequencyLoss(pred, target) (src/training/losses.py:16-35) with the 2-D FFT in shared memory: one
    CTA per image transforms pred - target, the spectrum-side arithmetic follows, and the backward runs the adjoint
    transform -- no cuFFT, no complex ATen glue.  pred / target: [B, C, H, W] fp32."""

    @staticmethod
    def forward(ctx, pred, target, alpha, eps, weight):
        B, Cc, H, W = pred.shape
        pred = pred.contiguous().float()
        target = target.contiguous().float()
        dev = pred.device
        spec = torch.empty((B, Cc, H, W // 2 + 1, 2), dtype=torch.float32, device=dev)
        loss = torch.empty((), dtype=torch.float32, device=dev)
        scale = torch.empty(Cc, dtype=torch.float32, device=dev)
        U = torch.empty(Cc, dtype=torch.float32, device=dev)
        ws = _ws(query("bvae_ffl_image_workspace", B, Cc, H, W), dev)
        with profiler.region("ffl_image_fwd", nbytes=2.0 * pred.numel() * 4 + spec.numel() * 4):
            call("bvae_ffl_image_fwd", _p(pred), _p(target), B, Cc, H, W, float(alpha), float(eps), float(weight), _p(spec),
                 _p(loss), _p(scale), _p(U), _p(ws), ws.numel(), _stream())
        ctx.save_for_backward(spec, scale, U)
        ctx.cfg = (tuple(pred.shape), float(alpha), float(eps), float(weight))
        return loss

    @staticmethod
    def backward(ctx, g_loss):
        spec, scale, U = ctx.saved_tensors
        shape, alpha, eps, weight = ctx.cfg
        B, Cc, H, W = shape
        d_pred = torch.empty(shape, dtype=torch.float32, device=spec.device)
        g_loss = g_loss.contiguous().float()
        ws = _ws(query("bvae_ffl_image_workspace", B, Cc, H, W), spec.device)
        with profiler.region("ffl_image_bwd", nbytes=spec.numel() * 4 + d_pred.numel() * 4):
            call("bvae_ffl_image_bwd", _p(spec), _p(scale), _p(U), _p(g_loss), weight, _p(d_pred), _p(ws), ws.numel(), B, Cc,
                 H, W, alpha, eps, _stream())
        return d_pred, None, None, None, None


def ffl_image_supported(H, W):
    return query("bvae_ffl_image_supported", int(H), int(W)) == 1


class LossCombine(torch.autograd.Function):
    """recon = base + lpips + ffl; total = recon + kl_side (src/models/beta_vae.py:216-217, 249-253) in one launch.
    Returns (total, recon, packed) where packed = [total, recon, kl_mean, recon_ffl, recon_base, recon_lpips]
    (non-differentiable; its first four entries are what a training loop logs)."""

    @staticmethod
    def forward(ctx, base, lp, ff, sc):
        out = torch.empty(6, dtype=torch.float32, device=base.device)
        call("bvae_loss_combine_fwd", _p(base), _p(lp), _p(ff), _p(sc), _p(out), _stream())
        ctx.has = (lp is not None, ff is not None)
        ctx.set_materialize_grads(False)
        packed = out.detach()
        ctx.mark_non_differentiable(packed)
        return out[0], out[1], packed

    @staticmethod
    def backward(ctx, g_total, g_rec, _g_packed):
        ref = g_total if g_total is not None else g_rec
        if ref is None:
            return None, None, None, None
        g = torch.empty(5, dtype=torch.float32, device=ref.device)
        call("bvae_loss_combine_bwd", _p(None if g_total is None else g_total.contiguous().float()),
             _p(None if g_rec is None else g_rec.contiguous().float()), _p(g), _stream())
        has_lp, has_ff = ctx.has
        return g[0], (g[0] if has_lp else None), (g[0] if has_ff else None), g[1:5]
