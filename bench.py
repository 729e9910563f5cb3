#!/usr/bin/env python
"""Benchmark of the Beta-VAE-SE training hot path on B200 (BASELINE.json metric:
train images/sec, fwd+bwd, at 1/2/4/8 GPUs, with the reference's CPU path timed beside it).

    python bench.py --gpus N --steps K --warmup W          # this repo's CUDA path
    python bench.py --impl reference --steps K --warmup W  # the reference's CPU path (oracle port)

A "step" is one full training step of the path on one batch: zero_grad, BetaVAE.loss() under
bf16 autocast (capacity objective C=30, gamma=1, focal-frequency term 0.5 -- the shipped default
config's first-epoch settings), backward, gradient all-reduce (N>1), global-norm clip 3.0 and
Adam.  Workload = beta-vae_b200/configs/beta_vae_se.yaml (128x128 grayscale, base 64, 4 blocks,
latent 64), per-GPU batch = training.batch_size (32), synthetic 4-class MRI-shaped slices,
random-init (Kaiming) weights.  Weak scaling: per-GPU batch fixed.
`value`: inputs resident in HBM, the whole step replayed as one CUDA graph.  `e2e`: the same step through
GraphedTrainStep with host buffers -- pinned uint8 image bytes copied host->device every step on the copy stream (one
step ahead, converted to ToTensor's floats on the device), every step's loss scalars copied to pinned host memory and
read on the host (one step behind).  `roofline`: the dominant kernel timed on the device with CUDA events.
Prints ONE JSON line (rank 0).
"""
import argparse
import json
import os
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
PKG = os.path.join(ROOT, "beta-vae_b200")
for p in (ROOT, os.path.join(PKG, "src"), PKG):
    if p not in sys.path:
        sys.path.insert(0, p)

import torch  # noqa: E402
import yaml  # noqa: E402

CONFIG = os.path.join(PKG, "configs", "beta_vae_se.yaml")
METRIC = "train images/sec (fwd+bwd)"
LOSS_KW = dict(beta=1.0, capacity=30.0, free_bits=0.0, capacity_weight=1.0)
WORKLOAD = "configs/beta_vae_se.yaml default Beta-VAE-SE train step (loss fwd + bwd + grad all-reduce + clip + Adam)"
OBJECTIVE = "capacity C=30 gamma=1, FFL 0.5, mse"


def weight_init(m):
    # what the reference's train loop applies (src/training/train.py:217-221)
    if isinstance(m, (torch.nn.Conv2d, torch.nn.ConvTranspose2d, torch.nn.Linear)):
        torch.nn.init.kaiming_normal_(m.weight, nonlinearity="relu")
        if m.bias is not None:
            torch.nn.init.zeros_(m.bias)


def load_cfg():
    with open(CONFIG) as f:
        return yaml.safe_load(f)


def measured_peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        with open(path) as f:
            d = json.load(f)
        return {"hbm_gbs": d["hbm_gbs"], "bf16_tflops": d["bf16_tflops"],
                "bf16_tflops_sustained": d.get("bf16_tflops_sustained", d["bf16_tflops"]), "source": "measured"}
    return {"hbm_gbs": 6650.0, "bf16_tflops": 1590.0, "bf16_tflops_sustained": 1400.0, "source": "fallback"}


# ----------------------------------------------------------------------------------------------
# reference arm: the reference's CPU implementation of the path (oracle port; /root/reference
# does not exist on the GPU box).  Also used, on a bounded sample, for `cpu_baseline`.
# ----------------------------------------------------------------------------------------------
def run_cpu_reference(cfg, batch, steps, warmup):
    from types import SimpleNamespace

    from oracle import beta_vae_oracle as O

    def ns(d):
        return SimpleNamespace(**{k: ns(v) for k, v in d.items()}) if isinstance(d, dict) else d

    threads = os.cpu_count() or 1
    torch.set_num_threads(threads)
    spec = O.spec_from_config(ns(cfg))
    sd = {k: v.requires_grad_(True) for k, v in O.init_params(spec, cfg["data"]["seed"]).items()}
    opt = torch.optim.Adam(list(sd.values()), lr=cfg["optimization"]["lr"])
    xs = [O.synthetic_batch(batch, spec.image_size, seed=s) for s in range(2)]
    g = torch.Generator().manual_seed(1)
    times = []
    for i in range(warmup + steps):
        eps = torch.randn(batch, spec.latent_dim, generator=g)
        t0 = time.perf_counter()
        opt.zero_grad(set_to_none=True)
        out = O.loss(sd, spec, xs[i % 2], eps, **LOSS_KW)
        out["total"].backward()
        torch.nn.utils.clip_grad_norm_(list(sd.values()), cfg["training"]["grad_clip"])
        opt.step()
        if i >= warmup:
            times.append(time.perf_counter() - t0)
    sec = sum(times) / len(times)
    return batch / sec, sec * 1e3, threads


def reference_main(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    cfg = load_cfg()
    batch = cfg["training"]["batch_size"]
    ips, ms, threads = run_cpu_reference(cfg, batch, max(1, args.steps), max(0, args.warmup))
    line = {
        "impl": "reference", "metric": METRIC, "value": ips, "unit": "images/s", "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": WORKLOAD, "image_size": cfg["data"]["image_size"], "per_gpu_batch": batch,
                   "global_batch": batch, "objective": OBJECTIVE,
                   "parallelism": "cpu: the reference path (oracle port) on all host cores, fp32"},
        "cpu_baseline": {"value": ips, "unit": "images/s", "cores": threads, "kind": "port",
                         "sample": f"{args.steps} full train steps at batch {batch} (oracle port of the reference "
                                   f"path, torch CPU fp32, {threads} threads)"},
        "e2e": {"value": ips, "unit": "images/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    print(json.dumps(line), flush=True)


# ----------------------------------------------------------------------------------------------
# clock sampling during the timed region
# ----------------------------------------------------------------------------------------------
class ClockSampler:
    def __init__(self, index):
        self.samples, self.reasons, self.max_mhz = [], set(), None
        self._stop = threading.Event()
        self._thr = None
        try:
            import pynvml
            pynvml.nvmlInit()
            self.nv = pynvml
            self.h = pynvml.nvmlDeviceGetHandleByIndex(index)
            self.max_mhz = pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM)
        except Exception:
            self.nv = None

    def _loop(self):
        nv = self.nv
        names = {
            getattr(nv, "nvmlClocksEventReasonHwSlowdown", 0x8): "hw_slowdown",
            getattr(nv, "nvmlClocksEventReasonHwThermalSlowdown", 0x40): "hw_thermal_slowdown",
            getattr(nv, "nvmlClocksEventReasonSwThermalSlowdown", 0x20): "sw_thermal_slowdown",
            getattr(nv, "nvmlClocksEventReasonSwPowerCap", 0x4): "sw_power_cap",
        }
        while not self._stop.is_set():
            try:
                self.samples.append(nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM))
                try:
                    r = nv.nvmlDeviceGetCurrentClocksEventReasons(self.h)
                except Exception:
                    r = nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.h)
                for bit, name in names.items():
                    if r & bit:
                        self.reasons.add(name)
            except Exception:
                pass
            self._stop.wait(0.004)

    def __enter__(self):
        if self.nv is not None:
            self._thr = threading.Thread(target=self._loop, daemon=True)
            self._thr.start()
        return self

    def __exit__(self, *a):
        self._stop.set()
        if self._thr is not None:
            self._thr.join()

    def result(self):
        s = sorted(self.samples)
        return {"sm_mhz": s[len(s) // 2] if s else None, "sm_max_mhz": self.max_mhz, "reasons": sorted(self.reasons),
                "samples": len(s)}


# ----------------------------------------------------------------------------------------------
# CUDA arm
# ----------------------------------------------------------------------------------------------
def cuda_main(args):
    import torch.distributed as dist

    import __graft_entry__ as entry
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device (the product has no CPU path); "
                         "use --impl reference for the CPU arm")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    if rank == 0:
        entry.build()
    if world > 1:
        dist.barrier()

    from bvae_rt import lib, ops, profiler
    from bvae_rt.data import synthetic_batch
    from bvae_rt.engine import GraphedTrainStep, TrainEngine
    from models.beta_vae import BetaVAE
    from utils.brain_tumor_utils.config_parser import freeze
    lib.load()
    assert lib.query("bvae_device_ok") == 1, "not a compute-capability-10.x device"

    cfg = load_cfg()
    batch = args.batch or cfg["training"]["batch_size"]
    size, latent = cfg["data"]["image_size"], cfg["model"]["latent_dim"]
    torch.manual_seed(cfg["data"]["seed"])
    model = BetaVAE(freeze(cfg))
    model.apply(weight_init)
    model = model.to(dev).train()
    engine = TrainEngine(model, lr=cfg["optimization"]["lr"], clip=cfg["training"]["grad_clip"])

    n_batches = 8  # rotate distinct device-resident batches so inputs are never L2-warm
    xs = [synthetic_batch(batch, size, seed=1000 * rank + i, device=dev) for i in range(n_batches)]
    gen = torch.Generator(device=dev).manual_seed(7 + rank)
    scal = torch.zeros(4, device=dev)

    def train_step_eager(x):
        engine.zero_grad()
        eps = torch.randn(batch, latent, device=dev, generator=gen)
        with torch.autocast("cuda", dtype=torch.bfloat16):
            out = model.loss(x, eps=eps, **LOSS_KW)
        out["total"].backward()
        engine.step()
        return out

    graphed = None if args.no_graph else GraphedTrainStep(engine, (batch, 1, size, size), seed=7 + rank)

    def train_step(x):
        # the whole step (zero_grad .. Adam) is one CUDA-graph replay; x may be a device or pinned host tensor
        return graphed(x, **LOSS_KW) if graphed is not None else train_step_eager(x)

    def sync_all():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
            torch.cuda.synchronize()

    def timed(fn, steps, finish=None):
        sync_all()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for i in range(steps):
            fn(i)
        if finish is not None:
            finish()                     # still inside the timed region
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1)
        if world > 1:
            t = torch.tensor([ms], device=dev)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            ms = float(t[0])
        sync_all()
        return ms

    warm = max(3, args.warmup)
    for i in range(warm):
        train_step(xs[i % n_batches])
    # ---- (1) device-resident throughput
    l0 = lib.query("bvae_launch_count")
    with ClockSampler(local) as clk:
        ms = timed(lambda i: train_step(xs[i % n_batches]), args.steps)
    launches = lib.query("bvae_launch_count") - l0
    if graphed is not None:
        launches = graphed.launches_per_step * args.steps    # kernels inside the replayed graph
    value = world * batch * args.steps / (ms / 1e3)

    # ---- (2) end to end through the public API with host buffers
    # the dataset's native format: uint8 image bytes (the reference decodes PNGs to bytes and ToTensor divides by 255);
    # they cross PCIe as bytes and become the same fp32 values on the device (bvae_u8_to_unit_f32, bit-exact)
    host = [(x * 255.0).round().to(torch.uint8).cpu().pin_memory() for x in xs]
    assert torch.equal(ops.u8_to_unit_f32(host[0].to(dev)), xs[0]), "uint8 round trip of the synthetic batch is not exact"
    xdev = torch.empty_like(xs[0])
    host_out = torch.empty(2, 4, dtype=torch.float32).pin_memory()     # double-buffered read-back of the loss scalars
    out_ready = [torch.cuda.Event(), torch.cuda.Event()]
    seen = [0]

    staged = [graphed.prefetch(host[0])] if graphed is not None else None

    def e2e_step(i):
        if graphed is not None:
            # this step's input was put in flight (pinned host -> device, copy stream) while the previous step ran;
            # the next step's H2D is enqueued right behind this step's launch: one H2D per step, overlapped
            out = train_step(staged[0])
            staged[0] = graphed.prefetch(host[(i + 1) % n_batches])
        else:
            ops.u8_to_unit_f32(host[i % n_batches].to(dev, non_blocking=True), xdev)
            out = train_step(xdev)
        k = seen[0] & 1
        if "scalars" in out:                                         # graphed step: packed inside the graph
            host_out[k].copy_(out["scalars"], non_blocking=True)     # D2H of the step's loss scalars
        else:
            scal[0], scal[1], scal[2], scal[3] = out["total"].detach(), out["recon"].detach(), \
                out["kl_mean"].detach(), out["recon_ffl"]
            host_out[k].copy_(scal, non_blocking=True)
        out_ready[k].record()
        # every step's scalars are read on the host, one step behind: the wait for step i-1's copy (already enqueued
        # ahead of step i's kernels) overlaps step i, so the device is never idle while Python catches up
        val = None
        if seen[0] > 0:
            out_ready[k ^ 1].synchronize()
            val = float(host_out[k ^ 1][0])
        seen[0] += 1
        return val

    def e2e_finish():
        # the last step's scalars are read too (every step's result reaches the host inside the timed region)
        k = (seen[0] - 1) & 1
        out_ready[k].synchronize()
        return float(host_out[k][0])

    for i in range(2):
        e2e_step(i)
    ms_e2e = timed(e2e_step, args.steps, finish=e2e_finish)
    e2e_value = world * batch * args.steps / (ms_e2e / 1e3)

    # ---- (3) rooflines.  (a) eager per-op CUDA-event pass: ranks the C-ABI calls (host launch gaps inflate the small
    #      ones, so it is only used for the `top_ops` table).  (b) the dominant kernel -- the stride-1 halo conv at
    #      the heaviest layer shape (dec3: 128x128, 64 -> 64 channels; with its dgrad twin 9 % of the step) -- and the
    #      two HBM-bound kernels north_star names are timed on the device: R launches captured in one CUDA graph,
    #      inputs rotating through buffers that together exceed the 126 MB L2, CUDA events around the replay.
    prof = profiler.start()
    for i in range(args.steps):
        train_step_eager(xs[i % n_batches])
    agg = profiler.stop().summary()
    peaks = measured_peaks()
    top = []
    if agg:
        total_ms = sum(a["ms"] for a in agg.values())
        ranked = sorted(agg.items(), key=lambda kv: -kv[1]["ms"])
        if rank == 0:
            print(f"# per-op CUDA-event profile (eager, includes launch gaps): {total_ms / args.steps:.3f} ms/step inside "
                  f"C-ABI calls", file=sys.stderr)
            for t, v in ranked:
                print(f"#  {t:42s} {v['ms'] / args.steps:8.4f} ms/step  {100 * v['ms'] / total_ms:5.1f}%  "
                      f"launches/call {v['launches'] / v['calls']:.0f}  "
                      f"{(v['flops'] / (v['ms'] * 1e-3) / 1e12) if v['flops'] else 0:8.1f} TFLOP/s  "
                      f"{(v['bytes'] / (v['ms'] * 1e-3) / 1e9) if v['bytes'] else 0:8.1f} GB/s", file=sys.stderr)
        for t, v in ranked[:12]:
            top.append({"op": t, "ms_per_step": v["ms"] / args.steps, "share": v["ms"] / total_ms,
                        "tflops": v["flops"] / (v["ms"] * 1e-3) / 1e12 if v["flops"] else None,
                        "gbs": v["bytes"] / (v["ms"] * 1e-3) / 1e9 if v["bytes"] else None})

    from bvae_rt import ops
    from bvae_rt.lib import BF16

    def device_us(fn, reps):
        """Average device time of one `fn(i)` launch group: `reps` of them captured in a CUDA graph, replayed 3x."""
        fn(0)
        torch.cuda.synchronize()
        g = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g):
            for i in range(reps):
                fn(i)
        g.replay()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        best = 1e30
        for _ in range(3):
            e0.record()
            g.replay()
            e1.record()
            torch.cuda.synchronize()
            best = min(best, e0.elapsed_time(e1) * 1e3 / reps)
        return best

    step_us = ms / args.steps * 1e3
    bf = torch.bfloat16
    # (b1) dominant kernel: conv_halo_tc_kernel<64, resident> on dec3 (fprop; the dgrad launch is the same kernel)
    Hc, Cc = size, cfg["model"]["base_channels"]
    xs_c = [torch.randn(batch, Hc, Hc, Cc, device=dev).to(bf) for _ in range(3)]
    w_c = torch.randn(Cc, Cc, 3, 3, device=dev) * 0.05
    wp_c = ops.pack_conv(w_c, BF16, False, ops.conv_weight_layout(Hc, Hc, Cc, Cc, 1, BF16, 0))
    b_c = torch.zeros(Cc, device=dev)
    conv_us = device_us(lambda i: ops.conv3x3_fprop(xs_c[i % 3], wp_c, b_c, Cc, 1, BF16), 9)
    conv_flops = 2.0 * batch * Hc * Hc * Cc * 9 * Cc
    ach = conv_flops / conv_us / 1e6
    roofline = {
        "kernel": f"conv_halo_tc_kernel<64,resident> dec3 conv3x3 {Hc}x{Hc}x{Cc}->{Cc} (tcgen05 implicit GEMM, fprop = dgrad kernel)",
        "bound": "tensor", "achieved": ach, "peak": peaks["bf16_tflops_sustained"], "unit": "TFLOP/s",
        "frac": ach / peaks["bf16_tflops_sustained"],
        # dram__bytes_read.sum + dram__bytes_write.sum of this launch at batch 32, profiles/r1_ncu_conv_halo.txt
        "traffic": 84.93e6 if batch == 32 else None,
        "algorithmic_bytes": 2.0 * batch * Hc * Hc * Cc * 2,
        "peak_source": peaks["source"] + " (sustained bf16, cuBLAS)", "us_per_launch": conv_us,
        "launches_per_step": 2, "share_of_step": 2 * conv_us / step_us,
        "note": "SS-mode tcgen05 fetches operands at ~64 B/clk: ceiling N/(N+128) = 33 % of the tensor pipe at N = 64 "
                "(ncu sm__pipe_tensor_cycles_active 33.4 %), see DESIGN.md 4.1",
    }
    del xs_c
    # (b2) HBM-bound kernels named by north_star: GroupNorm+ReLU+SE (cluster kernel, dec2 shape; multi-pass, dec3 shape)
    #      and the loss kernel (final conv + sigmoid + reconstruction loss)
    more = []

    def gn_case(Hs, C, nbuf):
        R = max(1, C // cfg["model"]["se_reduction_ratio"])
        ys = [torch.randn(batch, Hs, Hs, C, device=dev).to(bf) for _ in range(nbuf)]
        gam, bet = torch.ones(C, device=dev), torch.zeros(C, device=dev)
        w1, b1 = torch.randn(R, C, device=dev) * 0.1, torch.zeros(R, device=dev)
        w2, b2 = torch.randn(C, R, device=dev) * 0.1, torch.zeros(C, device=dev)
        us = device_us(lambda i: ops.gn_relu_se_fwd(ys[i % nbuf], gam, bet, w1, b1, w2, b2), 2 * nbuf)
        nbytes = 2.0 * ys[0].numel() * 2
        more.append({"kernel": f"gn_relu_se_fwd {Hs}x{Hs}x{C}", "bound": "hbm", "achieved": nbytes / us / 1e3,
                     "peak": peaks["hbm_gbs"], "unit": "GB/s", "frac": nbytes / us / 1e3 / peaks["hbm_gbs"],
                     "traffic": None, "us_per_launch": us})

    gn_case(size // 2, Cc, 8)
    gn_case(size, Cc, 3)
    hs_f = [torch.randn(batch, size, size, Cc, device=dev).to(bf) for _ in range(3)]
    w_f = torch.randn(1, Cc, 3, 3, device=dev) * 0.05
    b_f = torch.zeros(1, device=dev)
    tgt = torch.rand(batch, 1, size, size, device=dev)
    us = device_us(lambda i: ops.FinalConvSigmoidRecon.apply(hs_f[i % 3], w_f, b_f, tgt, 0), 6)
    nbytes = hs_f[0].numel() * 2 + 8.0 * tgt.numel()
    more.append({"kernel": "finalconv_sigmoid_recon_fwd (tcgen05 all-tap channel GEMM over the halo + stencil/sigmoid/loss epilogue)", "bound": "hbm",
                 "achieved": nbytes / us / 1e3, "peak": peaks["hbm_gbs"], "unit": "GB/s",
                 "frac": nbytes / us / 1e3 / peaks["hbm_gbs"], "traffic": 70.5e6 if batch == 32 else None,
                 "us_per_launch": us})
    del hs_f

    # ---- (4) CPU baseline (rank 0, N=1 only): bounded sample of the same workload
    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        ips, cms, threads = run_cpu_reference(cfg, batch if batch <= 32 else 32, steps=2, warmup=1)
        cpu = {"value": ips, "unit": "images/s", "cores": threads, "kind": "port",
               "sample": f"2 timed train steps (+1 warm-up) at batch {min(batch, 32)} of the same config: oracle port of "
                         f"the reference path, torch CPU fp32, {threads} threads, {cms:.0f} ms/step"}

    if rank == 0:
        act_mb = 2 * 2 * 1998848 * batch * 2 / 1e6  # y + out of every block, bf16, fwd (+ same again in bwd)
        line = {
            "metric": METRIC, "value": value, "unit": "images/s", "n_gpus": world, "steps": args.steps,
            "warmup": warm, "ms_per_step": ms / args.steps, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "bf16", "data": "synthetic",
            "config": {"workload": WORKLOAD,
                       "image_size": size, "per_gpu_batch": batch, "global_batch": batch * world,
                       "objective": OBJECTIVE,
                       "parallelism": f"dp{world}",
                       "launch": "eager" if graphed is None else "whole step captured in one CUDA graph",
                       "l2": f"no explicit flush: a step streams ~{act_mb:.0f} MB of activations plus 8 rotating "
                             f"input batches, above the 126 MB L2"},
            "clocks": clk.result(),
            "e2e": {"value": e2e_value, "unit": "images/s", "ms_per_step": ms_e2e / args.steps,
                    "h2d_bytes_per_step": host[0].numel() * host[0].element_size(), "d2h_bytes_per_step": 16,
                    "input": "pinned uint8 image bytes, prefetched on a copy stream, /255 on the device",
                    "readback": "4 loss scalars of every step -> pinned host memory, read one step behind"},
            "gpu_launches": int(launches),
            "roofline": roofline,
            "roofline_more": more,
            "cpu_baseline": cpu,
            "top_ops": top,
        }
        print(json.dumps(line), flush=True)
    if world > 1:
        # NCCL communicators referenced by a captured CUDA graph can stall destroy_process_group();
        # everything is flushed, so leave through a barrier and a hard exit
        graphed = None
        torch.cuda.synchronize()
        dist.barrier()
        sys.stdout.flush()
        sys.stderr.flush()
        os._exit(0)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=200)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--batch", type=int, default=0, help="per-GPU batch (default: training.batch_size of the config)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-graph", action="store_true", help="enqueue every kernel from Python instead of replaying a CUDA graph")
    args = ap.parse_args()
    if args.impl == "reference":
        reference_main(args)
    else:
        cuda_main(args)


if __name__ == "__main__":
    main()
