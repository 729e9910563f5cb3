#!/usr/bin/env python
"""Benchmark of the Beta-VAE-SE training hot path on B200 (BASELINE.json metric:
train images/sec, fwd+bwd, at 1/2/4/8 GPUs, with the reference's own implementation timed beside it).

    python bench.py --gpus N --steps K --warmup W                  # this repo's CUDA path (default workload)
    python bench.py --impl reference --steps K --warmup W          # the UNMODIFIED reference class on the host cores
    python bench.py --workload {default,overfit_capacity,overfit_capacity_wide128,inference}

A "step" is one full training step of the path on one batch: zero_grad, BetaVAE.loss() under
bf16 autocast, backward, gradient all-reduce (N>1), global-norm clip 3.0 and Adam.

Workloads (BASELINE.json configs):
  default                   configs/beta_vae_se.yaml (128x128 grayscale, base 64, 4 blocks, latent 64), per-GPU batch =
                            training.batch_size (32), capacity objective C=30 gamma=1 + focal-frequency term 0.5
                            (BASELINE configs 2 / 3: the line the driver records)
  overfit_capacity          configs/overfit_capacity.yaml as written (256x256, base 16, 5 blocks, latent 32, beta 4),
                            batch 8 unless --batch says otherwise (BASELINE config 4, HBM-bound as written)
  overfit_capacity_wide128  the same file with model.base_channels = 128: the tensor-core-bound variant SURVEY.md 8d asks for
  inference                 encode / sample_prior (decode) / forward(deterministic) at batch 1024 per GPU, default config
                            (BASELINE config 5); N GPUs = N independent replicas
Synthetic 4-class MRI-shaped slices, random-init (Kaiming) weights.  Weak scaling: per-GPU batch fixed.

`value`: inputs resident in HBM, the whole step replayed as one CUDA graph.  `e2e`: the same step through
GraphedTrainStep with HOST buffers -- pinned uint8 image bytes copied host->device every step on the copy stream (one
step ahead, converted to ToTensor's floats on the device), every step's loss scalars copied to pinned host memory and
read on the host (one step behind).  `roofline`: the kernel with the largest share of the step, timed on the device
with CUDA events in isolation -> divided by the BURST peak.  `cpu_baseline` / `--impl reference`: the unmodified
reference ``models.beta_vae.BetaVAE`` staged by oracle/make_ref.py (`kind: "reference"`; the oracle port only when the
staged copy is absent).  `gpu_eager_baseline`: the same reference class on this B200 through PyTorch eager
(cuDNN / cuBLAS / cuFFT) -- SURVEY.md 2b's "kernel to beat".  Prints ONE JSON line (rank 0).
"""
import argparse
import copy
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
PKG = os.path.join(ROOT, "beta-vae_b200")
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

import torch  # noqa: E402
import yaml  # noqa: E402

METRIC = "train images/sec (fwd+bwd)"

WORKLOADS = {
    "default": dict(
        yaml="beta_vae_se.yaml", override={}, batch=None, metric=METRIC,
        loss_kw=dict(beta=1.0, capacity=30.0, free_bits=0.0, capacity_weight=1.0),
        name="configs/beta_vae_se.yaml default Beta-VAE-SE train step (loss fwd + bwd + grad all-reduce + clip + Adam)",
        objective="capacity C=30 gamma=1, FFL 0.5, mse"),
    "overfit_capacity": dict(
        yaml="overfit_capacity.yaml", override={}, batch=None, metric=METRIC,
        loss_kw=dict(beta=4.0, capacity=None, free_bits=0.0, capacity_weight=None),
        name="configs/overfit_capacity.yaml train step as written (256x256, base 16, 5 blocks, latent 32, no decoder SE)",
        objective="beta=4 (first-epoch schedule value), mse, no FFL"),
    "overfit_capacity_wide128": dict(
        yaml="overfit_capacity.yaml", override={"model": {"base_channels": 128}}, batch=None, metric=METRIC,
        loss_kw=dict(beta=4.0, capacity=None, free_bits=0.0, capacity_weight=None),
        name="configs/overfit_capacity.yaml with model.base_channels=128 (tensor-core-bound variant, SURVEY.md 8d)",
        objective="beta=4, mse, no FFL"),
    "inference": dict(
        yaml="beta_vae_se.yaml", override={}, batch=1024, metric="inference images/sec (encode + decode)",
        loss_kw={},
        name="configs/beta_vae_se.yaml inference: forward(deterministic) = encode + decode, batch 1024 per GPU "
             "(src/inference encode.py / generate.py call sites)",
        objective="n/a (forward only)"),
}


def weight_init(m):
    # what the reference's train loop applies (src/training/train.py:217-221)
    if isinstance(m, (torch.nn.Conv2d, torch.nn.ConvTranspose2d, torch.nn.Linear)):
        torch.nn.init.kaiming_normal_(m.weight, nonlinearity="relu")
        if m.bias is not None:
            torch.nn.init.zeros_(m.bias)


def _merge(dst, src):
    for k, v in src.items():
        if isinstance(v, dict):
            _merge(dst.setdefault(k, {}), v)
        else:
            dst[k] = v
    return dst


def load_cfg(workload):
    w = WORKLOADS[workload]
    with open(os.path.join(PKG, "configs", w["yaml"])) as f:
        cfg = yaml.safe_load(f)
    return _merge(cfg, copy.deepcopy(w["override"]))


def config_key(workload, cfg, batch):
    """The `config` object: identical in both arms (what differs between them -- parallelism, launch mode, L2 policy,
    global batch -- lives in top-level keys)."""
    w = WORKLOADS[workload]
    return {"workload": w["name"], "image_size": cfg["data"]["image_size"], "per_gpu_batch": batch,
            "objective": w["objective"]}


def measured_peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        with open(path) as f:
            d = json.load(f)
        return {"hbm_gbs": d["hbm_gbs"], "bf16_tflops": d["bf16_tflops"],
                "bf16_tflops_sustained": d.get("bf16_tflops_sustained", d["bf16_tflops"]), "source": "measured"}
    return {"hbm_gbs": 6650.0, "bf16_tflops": 1590.0, "bf16_tflops_sustained": 1400.0, "source": "fallback"}


# ----------------------------------------------------------------------------------------------
# reference arm: the reference's own implementation of the path.  Preferred: the UNMODIFIED class staged by
# oracle/make_ref.py into oracle/_ref (kind "reference"); fallback: the oracle port (kind "port").
# /root/reference does not exist on the GPU box; oracle/_ref travels with the tree.
# ----------------------------------------------------------------------------------------------
def _ref_available():
    from oracle import make_ref
    return make_ref.available()


def run_reference(workload, device, amp, batch, steps, warmup):
    """Time `steps` train steps of the reference path.  Returns (images/s, ms/step, threads, kind)."""
    from oracle import beta_vae_oracle as O
    from oracle import make_ref
    w = WORKLOADS[workload]
    cfg = load_cfg(workload)
    kw = w["loss_kw"]
    threads = os.cpu_count() or 1
    torch.set_num_threads(threads)
    size, latent = cfg["data"]["image_size"], cfg["model"]["latent_dim"]
    dev = torch.device(device)
    xs = [O.synthetic_batch(batch, size, seed=s).to(dev) for s in range(2)]

    if make_ref.available():
        # the unmodified reference class, stepped the way its own loop does (src/training/train.py:313-330:
        # zero_grad, autocast(model.loss), GradScalerWrapper.backward = backward + unscale + clip + Adam)
        kind = "reference"
        import tempfile
        tmp = tempfile.mkdtemp(prefix="bvae_ref_cfg_")
        cfg_path = os.path.join(tmp, "cfg.yaml")
        with open(cfg_path, "w") as f:
            yaml.safe_dump(cfg, f)
        os.environ["CONFIG_PATH"] = cfg_path
        sys.path.insert(0, make_ref.src_dir())
        from models.beta_vae import BetaVAE as RefBetaVAE                  # noqa: E402  (oracle/_ref/src)
        from training.callbacks import GradScalerWrapper, get_optimizer    # noqa: E402
        assert os.path.realpath(sys.modules["models.beta_vae"].__file__).startswith(os.path.realpath(make_ref.DST)), \
            "reference arm imported a non-reference models.beta_vae"
        torch.manual_seed(cfg["data"]["seed"])
        model = RefBetaVAE()
        model.apply(weight_init)
        model.to(dev).train()
        opt = get_optimizer(model)
        scaler = GradScalerWrapper(enabled=(amp == "fp16"))
        scaler.set_params(model.parameters())
        clip = cfg["training"]["grad_clip"]

        def step(i):
            opt.zero_grad(set_to_none=True)
            if amp == "none":
                losses = model.loss(xs[i % 2], **kw)
            else:
                with torch.amp.autocast(dev.type, dtype=torch.float16 if amp == "fp16" else torch.bfloat16):
                    losses = model.loss(xs[i % 2], **kw)
            scaler.backward(losses["total"], opt, clip=clip if clip > 0 else 0.0)
            return losses["total"]
    else:
        kind = "port"
        from types import SimpleNamespace

        def ns(d):
            return SimpleNamespace(**{k: ns(v) for k, v in d.items()}) if isinstance(d, dict) else d

        spec = O.spec_from_config(ns(cfg))
        sd = {k: v.to(dev).requires_grad_(True) for k, v in O.init_params(spec, cfg["data"]["seed"]).items()}
        opt = torch.optim.Adam(list(sd.values()), lr=cfg["optimization"]["lr"])
        g = torch.Generator().manual_seed(1)

        def step(i):
            eps = torch.randn(batch, latent, generator=g).to(dev)
            opt.zero_grad(set_to_none=True)
            out = O.loss(sd, spec, xs[i % 2], eps, **kw)
            out["total"].backward()
            torch.nn.utils.clip_grad_norm_(list(sd.values()), cfg["training"]["grad_clip"])
            opt.step()
            return out["total"]

    def sync():
        if dev.type == "cuda":
            torch.cuda.synchronize()

    for i in range(warmup):
        step(i)
    sync()
    t0 = time.perf_counter()
    last = None
    for i in range(steps):
        last = step(i)
    if last is not None:
        float(last.detach())             # the loss reaches the host, as in the reference's loop (.item() per step)
    sync()
    sec = (time.perf_counter() - t0) / max(1, steps)
    return batch / sec, sec * 1e3, threads, kind


def reference_main(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    workload = args.workload
    if workload == "inference":
        print(json.dumps({"impl": "reference", "unavailable": "the inference workload has no reference arm in bench.py "
                          "(tests/test_gpu_inference.py checks it against the oracle)"}), flush=True)
        return
    cfg = load_cfg(workload)
    batch = args.batch or cfg["training"]["batch_size"]
    device, amp = args.ref_device, args.ref_amp
    if device == "cpu":
        amp = "none"
    ips, ms, threads, kind = run_reference(workload, device, amp, batch, max(1, args.steps), max(0, args.warmup))
    where = (f"cpu: {threads} host threads, fp32" if device == "cpu"
             else f"cuda eager (cuDNN/cuBLAS/cuFFT), autocast {amp}" + (" + GradScaler" if amp == "fp16" else ""))
    what = "unmodified reference class (oracle/_ref)" if kind == "reference" else "oracle port of the reference path"
    line = {
        "impl": "reference", "metric": WORKLOADS[workload]["metric"], "value": ips, "unit": "images/s",
        "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f32" if amp == "none" else amp, "data": "synthetic",
        "config": config_key(workload, cfg, batch),
        "global_batch": batch, "parallelism": where, "launch": "PyTorch eager", "l2": "n/a",
        "cpu_baseline": {"value": ips, "unit": "images/s", "cores": threads, "kind": kind,
                         "sample": f"{args.steps} full train steps at batch {batch}: {what}, {where}"},
        "e2e": {"value": ips, "unit": "images/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    print(json.dumps(line), flush=True)


def reference_subprocess(workload, device, amp, batch, steps, warmup, timeout=600):
    """Run the reference arm in its own interpreter (its `models` / `training` / `utils` packages must not meet the
    drop-in's in one process) and return its parsed JSON line, or a dict with `error`."""
    cmd = [sys.executable, os.path.abspath(__file__), "--impl", "reference", "--workload", workload, "--ref-device", device,
           "--ref-amp", amp, "--batch", str(batch), "--steps", str(steps), "--warmup", str(warmup)]
    env = {k: v for k, v in os.environ.items() if k not in ("RANK", "LOCAL_RANK", "WORLD_SIZE", "MASTER_ADDR", "MASTER_PORT")}
    try:
        out = subprocess.run(cmd, capture_output=True, text=True, timeout=timeout, env=env, cwd=ROOT)
        for ln in reversed(out.stdout.strip().splitlines()):
            if ln.startswith("{"):
                return json.loads(ln)
        return {"error": (out.stderr or out.stdout)[-400:]}
    except Exception as e:  # noqa: BLE001
        return {"error": repr(e)}


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


# ncu `dram__bytes_read.sum + dram__bytes_write.sum` per launch at the default workload's shapes (batch 32); sources
# under profiles/ (r1_ncu_conv_halo.txt, r1_ncu_wgrad_halo.txt, r1_ncu_gn_cluster.txt, r1_ncu_s2dgrad_finalconv.txt,
# refreshed by the r2_* files when a kernel changed)
NCU_TRAFFIC = {
    # dram__bytes_read.sum + dram__bytes_write.sum per launch, profiles/r2_ncu_step_kernels.txt (B = 32, in-step launches)
    "conv_fprop_dec3": 67.22e6 + 21.88e6,   # most of the 67 MB output is still in L2 when the kernel ends
    "conv_wgrad_dec3": 135.6e6 + 4.77e6,    # + 24.1 MB read by the split-K reduce (conv_wgrad_halo_reduce_kernel)
    "gn_fwd_64x64x64": 16.84e6 + 0.0,       # read = the tensor once; the write is still in L2 when the kernel ends
    "gn_fwd_128x128x64": (67.40e6 + 0.51e6) + (67.13e6 + 24.61e6),   # gn_pool_kernel + gn_scale_kernel (statistics: conv epilogue)
    "finalconv_fwd": 69.24e6 + 1.91e6,
}


# ----------------------------------------------------------------------------------------------
# CUDA arm
# ----------------------------------------------------------------------------------------------
def cuda_main(args):
    import torch.distributed as dist

    for p in (os.path.join(PKG, "src"), PKG):
        if p not in sys.path:
            sys.path.insert(0, p)
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

    workload = args.workload
    W = WORKLOADS[workload]
    cfg = load_cfg(workload)
    batch = args.batch or W["batch"] or cfg["training"]["batch_size"]
    size, latent = cfg["data"]["image_size"], cfg["model"]["latent_dim"]
    LOSS_KW = W["loss_kw"]
    torch.manual_seed(cfg["data"]["seed"])
    model = BetaVAE(freeze(cfg))
    model.apply(weight_init)

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
    n_batches = 8  # rotate distinct device-resident batches so inputs are never L2-warm
    xs = [synthetic_batch(batch, size, seed=1000 * rank + i, device=dev) for i in range(n_batches if batch <= 256 else 2)]
    n_batches = len(xs)
    # the dataset's native format: uint8 image bytes (the reference decodes PNGs to bytes and ToTensor divides by 255);
    # they cross PCIe as bytes and become the same fp32 values on the device (bvae_u8_to_unit_f32, bit-exact)
    host = [(x * 255.0).round().to(torch.uint8).cpu().pin_memory() for x in xs]

    def finish_line(line):
        if rank == 0:
            print(json.dumps(line), flush=True)
        if world > 1:
            # NCCL communicators referenced by a captured CUDA graph can stall destroy_process_group();
            # everything is flushed, so leave through a barrier and a hard exit
            torch.cuda.synchronize()
            dist.barrier()
            sys.stdout.flush()
            sys.stderr.flush()
            os._exit(0)

    # ==========================================================================================
    # inference workload (BASELINE config 5): N independent replicas, no collective
    # ==========================================================================================
    if workload == "inference":
        model = model.to(dev).eval()
        assert torch.equal(ops.u8_to_unit_f32(host[0].to(dev)), xs[0])
        res = {}
        with torch.no_grad(), torch.autocast("cuda", dtype=torch.bfloat16):
            fns = {"encode": lambda i: model.encode(xs[i % n_batches]),
                   "decode": lambda i: model.sample_prior(batch),
                   "forward": lambda i: model.forward(xs[i % n_batches], deterministic=True)}
            l0 = lib.query("bvae_launch_count")
            for name, fn in fns.items():
                for i in range(warm):
                    fn(i)
                with ClockSampler(local) as clk:
                    ms = timed(fn, args.steps)
                res[name] = {"ms": ms / args.steps, "images_per_s": world * batch * args.steps / (ms / 1e3)}
                if name == "forward":
                    clocks = clk.result()
            launches = lib.query("bvae_launch_count") - l0
            # end to end: pinned uint8 batch -> device, forward, reconstruction's first row + mu back to the host
            stage = [torch.empty_like(host[0], device=dev) for _ in range(2)]
            mu_host = torch.empty(batch, latent).pin_memory()

            def e2e(i):
                stage[i & 1].copy_(host[i % n_batches], non_blocking=True)
                recon, mu, _, _ = model.forward(stage[i & 1], deterministic=True)
                mu_host.copy_(mu, non_blocking=True)
                torch.cuda.current_stream().synchronize()
                return float(mu_host[0, 0])
            for i in range(2):
                e2e(i)
            ms_e2e = timed(e2e, args.steps)
        line = {
            "metric": W["metric"], "value": res["forward"]["images_per_s"], "unit": "images/s", "n_gpus": world,
            "steps": args.steps, "warmup": warm, "ms_per_step": res["forward"]["ms"], "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "bf16", "data": "synthetic",
            "config": config_key(workload, cfg, batch), "global_batch": batch * world,
            "parallelism": f"{world} independent replicas (no collective)", "launch": "eager (kernels enqueued from Python)",
            "l2": f"{n_batches} rotating input batches of {xs[0].numel() * 4 / 1e6:.0f} MB each",
            "clocks": clocks, "gpu_launches": int(launches),
            "e2e": {"value": world * batch * args.steps / (ms_e2e / 1e3), "unit": "images/s", "ms_per_step": ms_e2e / args.steps,
                    "h2d_bytes_per_step": host[0].numel(), "d2h_bytes_per_step": batch * latent * 4,
                    "input": "pinned uint8 image bytes", "readback": "mu [B, latent] to pinned host memory, read every step"},
            "parts": res,
        }
        finish_line(line)
        return

    # ==========================================================================================
    # training workloads
    # ==========================================================================================
    model = model.to(dev).train()
    engine = TrainEngine(model, lr=cfg["optimization"]["lr"], clip=cfg["training"]["grad_clip"])
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

    for i in range(warm):
        train_step(xs[i % n_batches])

    # ---- data-parallel consistency: the exchange sums exactly, and after warm-up every rank must hold bit-identical
    #      parameters
    dp_check = None
    if world > 1:
        torch.cuda.synchronize()
        # the gradient exchange on a known pattern: rank r contributes (r + 1) * (i mod 251) -> sum = W (W + 1) / 2 * (i mod 251)
        saved_g = engine.flat_g.clone()
        pat = (torch.arange(engine.numel, device=dev) % 251).float()
        engine.flat_g.copy_(pat * (rank + 1))
        if engine.dp_mode == "nvls":
            engine._nvls_all_reduce()
        else:
            dist.all_reduce(engine.flat_g)
        torch.cuda.synchronize()
        exact = bool(torch.equal(engine.flat_g, pat * (world * (world + 1) // 2)))
        ex = torch.tensor([1.0 if exact else 0.0], device=dev)
        dist.all_reduce(ex, op=dist.ReduceOp.MIN)
        # the exchange alone, device-timed (10 launches between two events, max over ranks)
        dist.barrier()
        torch.cuda.synchronize()
        ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        ev0.record()
        for _ in range(10):
            if engine.dp_mode == "nvls":
                engine._nvls_all_reduce()
            else:
                dist.all_reduce(engine.flat_g)
        ev1.record()
        torch.cuda.synchronize()
        xus = torch.tensor([ev0.elapsed_time(ev1) * 100.0], device=dev)
        dist.all_reduce(xus, op=dist.ReduceOp.MAX)
        exchange_us = float(xus[0])
        engine.flat_g.copy_(saved_g)
        del saved_g, pat
        torch.cuda.synchronize()
        ck = torch.stack([engine.flat_p.double().sum(), engine.flat_p.view(torch.int32).long().sum().double()])
        lo, hi = ck.clone(), ck.clone()
        dist.all_reduce(lo, op=dist.ReduceOp.MIN)
        dist.all_reduce(hi, op=dist.ReduceOp.MAX)
        dp_check = {"param_sum_min": float(lo[0]), "param_sum_max": float(hi[0]),
                    "param_bits_min": float(lo[1]), "param_bits_max": float(hi[1]),
                    "identical": bool(torch.equal(lo, hi)), "after_steps": warm,
                    "exchange": engine.dp_mode + (f"/{engine._nvls['algo']}" if engine._nvls is not None else ""),
                    "exchange_sums_exactly": bool(ex[0] > 0.5),
                    "exchange_timeouts": engine.nvls_status(), "exchange_us_isolated": exchange_us,
                    "exchange_bytes": engine.numel * 4,
                    "what": "sum and int32-bit-pattern sum of every rank's flat fp32 parameter buffer, all-reduce MIN vs MAX"}

    # ---- (1) device-resident throughput
    l0 = lib.query("bvae_launch_count")
    with ClockSampler(local) as clk:
        ms = timed(lambda i: train_step(xs[i % n_batches]), args.steps)
    launches = lib.query("bvae_launch_count") - l0
    if graphed is not None:
        launches = graphed.launches_per_step * args.steps    # kernels inside the replayed graph
    value = world * batch * args.steps / (ms / 1e3)

    # ---- (2) end to end through the public API with host buffers
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

    peaks = measured_peaks()
    step_us = ms / args.steps * 1e3
    line = {
        "metric": W["metric"], "value": value, "unit": "images/s", "n_gpus": world, "steps": args.steps,
        "warmup": warm, "ms_per_step": ms / args.steps, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "bf16", "data": "synthetic",
        "config": config_key(workload, cfg, batch),
        "global_batch": batch * world, "parallelism": f"dp{world}" + (f" ({engine.dp_mode}{'/' + engine._nvls['algo'] if engine._nvls is not None else ''} gradient exchange)" if world > 1 else ""),
        "launch": "eager" if graphed is None else "whole step captured in one CUDA graph",
        "l2": f"no explicit flush: {n_batches} rotating input batches and every step streams its activations "
              f"(y + out of every block, fwd and bwd) through HBM",
        "clocks": clk.result(),
        "e2e": {"value": e2e_value, "unit": "images/s", "ms_per_step": ms_e2e / args.steps,
                "h2d_bytes_per_step": host[0].numel() * host[0].element_size(), "d2h_bytes_per_step": 16,
                "input": "pinned uint8 image bytes, prefetched on a copy stream, /255 on the device",
                "readback": "4 loss scalars of every step -> pinned host memory, read one step behind"},
        "gpu_launches": int(launches),
    }
    if dp_check is not None:
        line["dp_check"] = dp_check

    if workload != "default":
        finish_line(line)
        return

    # ---- (3) rooflines (default workload).  (a) eager per-op CUDA-event pass: ranks the C-ABI calls (host launch gaps
    #      inflate the small ones, so it is only used for the `top_ops` table).  (b) the heaviest kernels -- the dec3
    #      weight-gradient and forward convolutions -- and the two HBM-bound kernels north_star names are timed on the
    #      device IN ISOLATION: R launches captured in one CUDA graph, inputs rotating through buffers that together
    #      exceed the 126 MB L2, CUDA events around the replay.  Isolated kernels are divided by the BURST peak.
    prof = profiler.start()
    for i in range(min(args.steps, 20)):
        train_step_eager(xs[i % n_batches])
    agg = profiler.stop().summary()
    nprof = min(args.steps, 20)
    top = []
    if agg:
        total_ms = sum(a["ms"] for a in agg.values())
        ranked = sorted(agg.items(), key=lambda kv: -kv[1]["ms"])
        if rank == 0:
            print(f"# per-op CUDA-event profile (eager, includes launch gaps): {total_ms / nprof:.3f} ms/step inside "
                  f"C-ABI calls", file=sys.stderr)
            for t, v in ranked:
                print(f"#  {t:42s} {v['ms'] / nprof:8.4f} ms/step  {100 * v['ms'] / total_ms:5.1f}%  "
                      f"launches/call {v['launches'] / v['calls']:.0f}  "
                      f"{(v['flops'] / (v['ms'] * 1e-3) / 1e12) if v['flops'] else 0:8.1f} TFLOP/s  "
                      f"{(v['bytes'] / (v['ms'] * 1e-3) / 1e9) if v['bytes'] else 0:8.1f} GB/s", file=sys.stderr)
        for t, v in ranked[:12]:
            top.append({"op": t, "ms_per_step": v["ms"] / nprof, "share": v["ms"] / total_ms,
                        "tflops": v["flops"] / (v["ms"] * 1e-3) / 1e12 if v["flops"] else None,
                        "gbs": v["bytes"] / (v["ms"] * 1e-3) / 1e9 if v["bytes"] else None})

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

    bf = torch.bfloat16
    Hc, Cc = size, cfg["model"]["base_channels"]
    burst = peaks["bf16_tflops"]
    conv_flops = 2.0 * batch * Hc * Hc * Cc * 9 * Cc
    alg_bytes = 2.0 * batch * Hc * Hc * Cc * 2
    xs_c = [torch.randn(batch, Hc, Hc, Cc, device=dev).to(bf) for _ in range(3)]
    dys_c = [torch.randn(batch, Hc, Hc, Cc, device=dev).to(bf) for _ in range(3)]
    w_c = torch.randn(Cc, Cc, 3, 3, device=dev) * 0.05
    wp_c = ops.pack_conv(w_c, BF16, False, ops.conv_weight_layout(Hc, Hc, Cc, Cc, 1, BF16, 0))
    b_c = torch.zeros(Cc, device=dev)
    dw_c, db_c = torch.empty_like(w_c), torch.empty_like(b_c)
    fprop_us = device_us(lambda i: ops.conv3x3_fprop(xs_c[i % 3], wp_c, b_c, Cc, 1, BF16), 9)
    wgrad_us = device_us(lambda i: ops.conv3x3_wgrad(xs_c[i % 3], dys_c[i % 3], 1, dw=dw_c, db=db_c), 9)
    del xs_c, dys_c

    def conv_entry(kernel, us, launches_per_step, traffic, note):
        ach = conv_flops / us / 1e6
        return {"kernel": kernel, "bound": "tensor", "achieved": ach, "peak": burst, "unit": "TFLOP/s", "frac": ach / burst,
                "traffic": traffic if batch == 32 else None, "algorithmic_bytes": alg_bytes,
                "peak_source": peaks["source"] + " (burst bf16, cuBLAS: the kernel is timed in isolation)",
                "us_per_launch": us, "launches_per_step": launches_per_step, "share_of_step": launches_per_step * us / step_us,
                "frac_of_sustained": ach / peaks["bf16_tflops_sustained"], "note": note}

    wgrad_entry = conv_entry(
        f"conv3x3 wgrad dec3 {Hc}x{Hc}x{Cc}->{Cc} (tcgen05, MN-major halo windows, split-K partials + reduce included)",
        wgrad_us, 1, NCU_TRAFFIC["conv_wgrad_dec3"],
        "the weight-gradient kernels are the largest kernel group of the step (dec3 + dec2 + dec1 + dec0 + encoder wgrads)")
    fprop_entry = conv_entry(
        f"conv3x3 fprop dec3 {Hc}x{Hc}x{Cc}->{Cc} (tcgen05 implicit GEMM over a shared-memory halo; dgrad = same kernel)",
        fprop_us, 2, NCU_TRAFFIC["conv_fprop_dec3"], "fprop + dgrad of dec3 run this kernel twice per step")
    # `roofline` = the weight-gradient kernel (conv_wgrad_halo_kernel has the largest share of the step over its six
    # launches); the forward / input-gradient kernel is the first entry of `roofline_more`
    roofline = wgrad_entry
    more = [fprop_entry]

    # (b2) HBM-bound kernels named by north_star: GroupNorm+ReLU+SE and the loss kernel
    def gn_case(Hs, C, nbuf, key, conv_stats=False):
        """``conv_stats``: the tensor comes out of the 3x3 convolution with its GroupNorm statistic partials, as in the
        step (the 2 MB samples of dec3 take the streaming passes, whose statistics ride in the conv epilogue)."""
        R = max(1, C // cfg["model"]["se_reduction_ratio"])
        gam, bet = torch.ones(C, device=dev), torch.zeros(C, device=dev)
        w1, b1 = torch.randn(R, C, device=dev) * 0.1, torch.zeros(R, device=dev)
        w2, b2 = torch.randn(C, R, device=dev) * 0.1, torch.zeros(C, device=dev)
        if conv_stats:
            wc = torch.randn(C, C, 3, 3, device=dev) * 0.05
            wp = ops.pack_conv(wc, BF16, False, ops.conv_weight_layout(Hs, Hs, C, C, 1, BF16, 0))
            pairs = [ops.conv3x3_fprop(torch.randn(batch, Hs, Hs, C, device=dev).to(bf), wp, torch.zeros(C, device=dev), C, 1,
                                       BF16, want_stats=True) for _ in range(nbuf)]
            ys, sps = [p[0] for p in pairs], [p[1] for p in pairs]
        else:
            ys = [torch.randn(batch, Hs, Hs, C, device=dev).to(bf) for _ in range(nbuf)]
            sps = [None] * nbuf
        us = device_us(lambda i: ops.gn_relu_se_fwd(ys[i % nbuf], gam, bet, w1, b1, w2, b2, stat_part=sps[i % nbuf]), 2 * nbuf)
        nbytes = 2.0 * ys[0].numel() * 2
        more.append({"kernel": f"gn_relu_se_fwd {Hs}x{Hs}x{C}" + (" (statistics from the conv epilogue)" if sps[0] is not None else ""),
                     "bound": "hbm", "achieved": nbytes / us / 1e3,
                     "peak": peaks["hbm_gbs"], "unit": "GB/s", "frac": nbytes / us / 1e3 / peaks["hbm_gbs"],
                     "traffic": NCU_TRAFFIC.get(key) if batch == 32 else None, "algorithmic_bytes": nbytes,
                     "us_per_launch": us})

    gn_case(size // 2, Cc, 8, "gn_fwd_64x64x64")
    gn_case(size, Cc, 3, "gn_fwd_128x128x64", conv_stats=True)
    hs_f = [torch.randn(batch, size, size, Cc, device=dev).to(bf) for _ in range(3)]
    w_f = torch.randn(1, Cc, 3, 3, device=dev) * 0.05
    b_f = torch.zeros(1, device=dev)
    tgt = torch.rand(batch, 1, size, size, device=dev)
    us = device_us(lambda i: ops.FinalConvSigmoidRecon.apply(hs_f[i % 3], w_f, b_f, tgt, 0), 6)
    nbytes = hs_f[0].numel() * 2 + 8.0 * tgt.numel()
    more.append({"kernel": "finalconv_sigmoid_recon_fwd (tcgen05 all-tap channel GEMM over the halo + stencil/sigmoid/loss epilogue)",
                 "bound": "hbm", "achieved": nbytes / us / 1e3, "peak": peaks["hbm_gbs"], "unit": "GB/s",
                 "frac": nbytes / us / 1e3 / peaks["hbm_gbs"], "traffic": NCU_TRAFFIC["finalconv_fwd"] if batch == 32 else None,
                 "algorithmic_bytes": nbytes, "us_per_launch": us})
    del hs_f

    # ---- (4) baselines (rank 0, N=1 only), each in its own interpreter: the reference on the host cores (bounded sample)
    #      and the reference through PyTorch eager on this GPU
    cpu = gpu_eager = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        engine.close()
        del graphed
        torch.cuda.synchronize()
        cb = min(batch, 32)
        r = reference_subprocess(workload, "cpu", "none", cb, steps=3, warmup=1)
        if "error" not in r and "cpu_baseline" in r:
            cpu = r["cpu_baseline"]
            cpu["sample"] = (f"3 timed train steps (+1 warm-up) at batch {cb} of the same config, "
                             f"{r['ms_per_step']:.0f} ms/step: " + cpu["sample"].split(": ", 1)[-1])
        else:
            cpu = {"value": None, "unit": "images/s", "cores": os.cpu_count(), "kind": "unavailable", "sample": str(r)[:300]}
        gpu_eager = {}
        for amp in ("fp16", "bf16"):
            r = reference_subprocess(workload, "cuda", amp, batch, steps=30, warmup=10)
            if "error" not in r and "value" in r:
                gpu_eager[amp] = {"value": r["value"], "unit": "images/s", "ms_per_step": r["ms_per_step"],
                                  "kind": r["cpu_baseline"]["kind"], "what": r["parallelism"]}
            else:
                gpu_eager[amp] = {"value": None, "error": str(r)[:300]}
        gpu_eager["note"] = ("the reference class on this same B200 through PyTorch eager; fp16 + GradScaler is what the "
                             "reference's train.py runs (torch.amp.autocast('cuda') default), bf16 is the like-for-like dtype")

    line.update({"roofline": roofline, "roofline_more": more, "cpu_baseline": cpu, "gpu_eager_baseline": gpu_eager,
                 "top_ops": top})
    finish_line(line)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=200)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--workload", default="default", choices=sorted(WORKLOADS))
    ap.add_argument("--batch", type=int, default=0, help="per-GPU batch (default: the workload's / config's batch size)")
    ap.add_argument("--no-cpu-baseline", action="store_true", help="skip the cpu_baseline and gpu_eager_baseline legs")
    ap.add_argument("--no-graph", action="store_true", help="enqueue every kernel from Python instead of replaying a CUDA graph")
    ap.add_argument("--ref-device", default="cpu", choices=["cpu", "cuda"], help="--impl reference: where the reference runs")
    ap.add_argument("--ref-amp", default="none", choices=["none", "fp16", "bf16"], help="--impl reference on cuda: autocast dtype")
    args = ap.parse_args()
    if args.impl == "reference":
        reference_main(args)
    else:
        cuda_main(args)


if __name__ == "__main__":
    main()
