"""Training engine for the B200 path: flat fp32 parameter / gradient / Adam-state buffers,
bucketed gradient all-reduce overlapped with backward (batch data parallel, one process per
GPU), and the fused global-norm-clip + Adam step.

What it stands in for in the reference: ``GradScalerWrapper.backward`` (backward, then
``clip_grad_norm_`` + ``optimizer.step()``; src/training/callbacks.py:18-30) and
``get_optimizer`` (torch.optim.Adam / AdamW; callbacks.py:92-102).  The reference has no
multi-GPU code; data parallelism is this repo's addition:

  * every layer is per-sample (GroupNorm(1,C) has per-sample statistics), so the batch is
    partitioned across ranks with no activation exchange;
  * one NCCL all-reduce (average) of the gradients per step, in four buckets laid out in
    backward-ready order -- decoder+final conv, fc_dec, fc_mu+fc_logvar, encoder -- each
    launched from a post-accumulate-grad hook as soon as its last gradient lands, so it runs
    on NCCL's stream under the rest of backward;
  * wgrad kernels write straight into the bucket memory (``ops.GradArena``), so there is no
    gather copy; clip + Adam then run once over the flat buffers (3 launches for the norm and
    the update of all 72 tensors).

Batch-coupled loss terms (capacity |mean KL - C|, free bits, FFL's batch-mean weight) use the
LOCAL batch on each rank, i.e. DistributedDataParallel semantics.
"""
from __future__ import annotations

import math
import os
from typing import List, Optional

import torch
import torch.distributed as dist

from . import ops
from .lib import call, query


def backward_ready_buckets(model) -> List[List[str]]:
    """Parameter names grouped into the four buckets of SURVEY.md 8e, each in the order the
    gradients become available during backward."""
    names = [n for n, _ in model.named_parameters()]

    def take(prefix):
        return [n for n in names if n.startswith(prefix)]

    nb = len(model.encoder)
    b1 = take("final_conv.")
    for i in reversed(range(nb)):
        b1 += take(f"decoder_blocks.{i}.")
    b2 = take("fc_dec.")
    # the deepest encoder blocks hold most of the encoder's parameters (enc3's conv alone 72 % at the default config)
    # and finish early in the encoder's backward: they ride with the heads bucket, so the last bucket -- the one whose
    # all-reduce nothing can hide -- is only the two shallow blocks
    early = max(0, nb - 2)
    b3 = take("fc_logvar.") + take("fc_mu.")
    for i in reversed(range(nb - early, nb)):
        b3 += take(f"encoder.{i}.")
    b4 = []
    for i in reversed(range(nb - early)):
        b4 += take(f"encoder.{i}.")
    got = b1 + b2 + b3 + b4
    assert sorted(got) == sorted(names), "bucket plan does not cover the parameter list"
    return [b for b in (b1, b2, b3, b4) if b]


class TrainEngine:
    def __init__(self, model, lr: float, betas=(0.9, 0.999), eps: float = 1e-8, weight_decay: float = 0.0,
                 decoupled_wd: bool = False, clip: float = 0.0, process_group=None, reduce_dtype=torch.float32,
                 dp_mode: Optional[str] = None):
        """``dp_mode``: how the gradient exchange is issued when world > 1 --
        "nvls": this repo's own all-reduce kernel (csrc/allreduce_nvls.cu) on the flat gradient buffer, which then lives in
                symmetric memory bound to an NVSwitch multicast object (in-switch fp32 reduction): ONE launch after
                backward, 138 us for 38 MB on 8 x B200 against NCCL's 184 us (default on CUDA when the group supports
                multicast, else "single").  BVAE_NVLS_OVERLAP=1 issues it per backward-ready bucket on a side stream
                instead -- measured SLOWER (1.70 vs 1.60 ms at 2 GPUs): a forked CUDA graph loses more in kernel-to-kernel
                latency than the hidden exchange gains, as round 1's side-stream experiment found;
        "single": ONE NCCL all-reduce of the whole flat gradient buffer after backward;
        "buckets": the four backward-ready buckets, each launched from a post-accumulate hook under the rest of backward.
        Measured on 8 x B200 (profiles/): the persistent convolution kernels own every SM, so NCCL's CTAs only run at
        kernel boundaries -- the bucketed launches cost more in disturbed kernels and launch latency than the exposed
        ~100 us of one 38 MB NVLS all-reduce.  BVAE_DP_MODE overrides."""
        import os
        self.dp_mode = os.environ.get("BVAE_DP_MODE", dp_mode or "auto")
        if self.dp_mode not in ("auto", "nvls", "single", "buckets"):
            raise ValueError(f"dp_mode '{self.dp_mode}' (auto | nvls | single | buckets)")
        self._nvls = None
        self.model = model
        self.lr, self.betas, self.eps = float(lr), tuple(betas), float(eps)
        self.weight_decay, self.decoupled_wd, self.clip = float(weight_decay), bool(decoupled_wd), float(clip)
        self.step_count = 0
        self.pg = process_group
        self.world = dist.get_world_size(process_group) if (dist.is_available() and dist.is_initialized()) else 1
        if os.environ.get("BVAE_DP_REDUCE_DTYPE", "") == "bf16":      # opt-in: halves the bytes, rounds the summands
            reduce_dtype = torch.bfloat16
        self.reduce_dtype = reduce_dtype
        named = dict(model.named_parameters())
        self._named = named
        dev = next(model.parameters()).device
        self.device = dev
        self.bucket_names = backward_ready_buckets(model)
        order = [n for b in self.bucket_names for n in b]
        sizes = [named[n].numel() for n in order]
        # 16-byte aligned slots so every view can be read with vector loads
        offs, off = [], 0
        for s in sizes:
            offs.append(off)
            off += (s + 3) // 4 * 4
        if self.world > 1:
            off = (off + 4 * self.world * 128 - 1) // (4 * self.world * 128) * (4 * self.world * 128)   # even rank slices
        self.numel = off
        self.flat_p = torch.zeros(off, dtype=torch.float32, device=dev)
        self.flat_g = self._alloc_grad_buffer(off, dev)
        self.flat_m = torch.zeros(off, dtype=torch.float32, device=dev)
        self.flat_v = torch.zeros(off, dtype=torch.float32, device=dev)
        self.slots = {}
        with torch.no_grad():
            for n, o, s in zip(order, offs, sizes):
                p = named[n]
                pv = self.flat_p[o:o + s].view_as(p)
                pv.copy_(p.data)
                p.data = pv                                   # parameters now live in the flat buffer
                self.slots[n] = (o, s, self.flat_g[o:o + s].view_as(p))
        self.bucket_ranges = []
        for b in self.bucket_names:
            lo = self.slots[b[0]][0]
            last = self.slots[b[-1]]
            self.bucket_ranges.append((lo, last[0] + (last[1] + 3) // 4 * 4))
        self._pending = [0] * len(self.bucket_names)
        self._works = []
        self._bucket_of = {n: i for i, b in enumerate(self.bucket_names) for n in b}
        self._sumsq = torch.zeros(1, dtype=torch.float32, device=dev)
        self._hyper = torch.zeros(4, dtype=torch.float32, device=dev)   # lr, 1-b1^t, sqrt(1-b2^t)
        self._hyper_host = None      # pinned staging ring for _hyper (advance())
        self._hyper_done = None
        self._ws = torch.empty(query("bvae_adam_workspace", self.numel), dtype=torch.uint8, device=dev)
        self._hooks = []
        arena = ops.GradArena()
        for n, p in named.items():
            arena.register(p, self.slots[n][2])
            self._hooks.append(p.register_post_accumulate_grad_hook(self._make_hook(n)))
        self.arena = arena
        ops.set_grad_arena(arena)
        # the GroupNorm / SE parameter gradients are reduced in one launch at the END of backward (ops.flush_gn_param_jobs)
        # -- unless gradient buckets leave from hooks during backward, which would ship them unreduced
        early = self.world > 1 and (self.dp_mode == "buckets" or (self._nvls is not None and self._nvls["overlap"]))
        ops.set_gn_param_deferral(not early)
        if hasattr(model, "invalidate_packs"):
            model.invalidate_packs()
        self.last_grad_norm = None

    # ------------------------------------------------------------------ NVLS exchange (symmetric gradient buffer)
    # Exchange kernel and its grid, the same on every rank.  Measured (profiles/r2_scale_n{2,4,8}_{mc,p2p}*.json, 38.3 MB):
    # peer loads / stores beat the multicast form at 2 ranks (89 vs 138 us) and 4 ranks (119 vs 133 us); at 8 ranks both
    # take 137-139 us, and the in-switch reduction is kept there.  BVAE_NVLS_ALGO = mc | p2p and BVAE_NVLS_BLOCKS override.
    NVLS_BLOCKS = 64
    NVLS_ALGO = "auto"

    def _alloc_grad_buffer(self, n, dev):
        """The flat gradient buffer.  With world > 1 on CUDA it is allocated as symmetric memory and bound to the group's
        NVSwitch multicast object (torch.distributed._symmetric_memory does the VMM / fabric-handle plumbing); the
        exchange is then csrc/allreduce_nvls.cu.  Anything missing (CPU / gloo, no multicast) -> plain memory + NCCL."""
        want = self.dp_mode in ("auto", "nvls")
        if not (want and self.world > 1 and dev.type == "cuda"):
            if self.dp_mode in ("auto", "nvls"):
                if self.dp_mode == "nvls" and self.world > 1:
                    raise RuntimeError("dp_mode 'nvls' needs CUDA tensors")
                self.dp_mode = "single"
            return torch.zeros(n, dtype=torch.float32, device=dev)
        try:
            import torch.distributed._symmetric_memory as symm
            group = self.pg if self.pg is not None else dist.group.WORLD
            gname = group.group_name
            if hasattr(symm, "enable_symm_mem_for_group"):
                try:
                    symm.enable_symm_mem_for_group(gname)
                except Exception:  # noqa: BLE001  (newer torch enables groups implicitly)
                    pass
            g = symm.empty(n, dtype=torch.float32, device=dev)
            g.zero_()
            hdl = symm.rendezvous(g, group=gname)
            if not hdl.multicast_ptr:
                raise RuntimeError("the group has no NVLS multicast support")
            algo = os.environ.get("BVAE_NVLS_ALGO", self.NVLS_ALGO)
            if algo == "auto":
                algo = "p2p" if self.world in (2, 4) else "mc"
            if algo == "p2p" and self.world not in (2, 4, 8):
                algo = "mc"
            self.NVLS_BLOCKS = int(os.environ.get("BVAE_NVLS_BLOCKS", "128" if algo == "p2p" else "64"))
            nflags = query("bvae_allreduce_nvls_flag_count", self.NVLS_BLOCKS, self.world)
            flags = symm.empty(nflags, dtype=torch.int32, device=dev)
            flags.zero_()
            fh = symm.rendezvous(flags, group=gname)
            state = torch.zeros(query("bvae_allreduce_nvls_state_bytes"), dtype=torch.uint8, device=dev)
            torch.cuda.synchronize(dev)
            dist.barrier(group=self.pg)                          # every rank's flags are zero before the first exchange
            # algo "mc" = in-switch reduction through the multicast object (allreduce_nvls_kernel), "p2p" = the same exchange
            # through plain peer mappings (allreduce_p2p_kernel); both on this symmetric buffer
            self._nvls = {"hdl": hdl, "fh": fh, "flags": flags, "state": state, "mc": int(hdl.multicast_ptr),
                          "buf_ptrs": int(hdl.buffer_ptrs_dev), "algo": algo,
                          "flag_ptrs": int(fh.buffer_ptrs_dev), "n": n, "stream": torch.cuda.Stream(device=dev),
                          "overlap": os.environ.get("BVAE_NVLS_OVERLAP", "0") == "1", "launched": 0}
            self.dp_mode = "nvls"
            return g
        except Exception as e:  # noqa: BLE001
            if self.dp_mode == "nvls":
                raise
            import warnings
            warnings.warn(f"NVLS gradient exchange unavailable ({e}); using NCCL")
            self.dp_mode = "single"
            return torch.zeros(n, dtype=torch.float32, device=dev)

    def _nvls_all_reduce(self, lo=0, hi=None):
        """Sum floats [lo, hi) of the gradient buffer over the ranks, on the current stream."""
        nv = self._nvls
        hi = nv["n"] if hi is None else hi
        st = torch.cuda.current_stream().cuda_stream
        if nv["algo"] == "p2p":
            call("bvae_allreduce_p2p", nv["buf_ptrs"], lo, hi - lo, nv["flag_ptrs"], nv["state"].data_ptr(),
                 dist.get_rank(self.pg), self.world, self.NVLS_BLOCKS, st)
        else:
            call("bvae_allreduce_nvls", nv["mc"] + 4 * lo, hi - lo, nv["flag_ptrs"], nv["state"].data_ptr(),
                 dist.get_rank(self.pg), self.world, self.NVLS_BLOCKS, st)

    def _launch_reduce_nvls(self, bi):
        """Bucket ``bi`` is complete on the current stream: exchange it on the side stream (fork; finish_reduce joins)."""
        nv = self._nvls
        cur = torch.cuda.current_stream()
        ready = torch.cuda.Event()
        ready.record(cur)
        side = nv["stream"]
        side.wait_event(ready)
        lo, hi = self.bucket_ranges[bi]
        with torch.cuda.stream(side):
            self._nvls_all_reduce(lo, hi)
        nv["launched"] += 1

    def nvls_status(self) -> int:
        """0, or 1 when an exchange timed out waiting for a peer (reading it syncs)."""
        return 0 if self._nvls is None else int(self._nvls["state"][8:12].view(torch.int32)[0])

    # ------------------------------------------------------------------ backward-side plumbing
    def _make_hook(self, name):
        view = self.slots[name][2]
        bi = self._bucket_of[name]

        def hook(p):
            g = p.grad
            if g is not None and g.data_ptr() != view.data_ptr():
                view.copy_(g)                                  # autograd cloned or accumulated elsewhere
                p.grad = view
            self._pending[bi] -= 1
            if self._pending[bi] == 0 and self.world > 1:
                if self.dp_mode == "buckets":
                    self._launch_reduce(bi)
                elif self.dp_mode == "nvls" and self._nvls["overlap"]:
                    self._launch_reduce_nvls(bi)
        return hook

    def _launch_reduce(self, bi):
        import os
        if str(bi) in os.environ.get("BVAE_DP_SKIP_BUCKETS", "").split(","):   # dev knob: measure exposed communication
            return
        lo, hi = self.bucket_ranges[bi]
        buf = self.flat_g[lo:hi]
        if self.reduce_dtype == torch.float32:
            w = dist.all_reduce(buf, op=dist.ReduceOp.SUM, group=self.pg, async_op=True)
            self._works.append((w, None, buf))
        else:
            low = buf.to(self.reduce_dtype)
            w = dist.all_reduce(low, op=dist.ReduceOp.SUM, group=self.pg, async_op=True)
            self._works.append((w, low, buf))

    def zero_grad(self):
        """Drop gradients (set_to_none semantics) and arm the bucket counters for the next backward."""
        for p in self.model.parameters():
            p.grad = None
        self._pending = [len(b) for b in self.bucket_names]
        self._works = []
        if self._nvls is not None:
            self._nvls["launched"] = 0

    def finish_reduce(self):
        """Wait for the bucket all-reduces (stream-side) and restore fp32 gradients; in "single" mode issue the one
        all-reduce of the whole gradient buffer here (backward has been enqueued in full at this point)."""
        if self.world > 1 and self.dp_mode in ("single", "nvls"):
            import os
            if os.environ.get("BVAE_DP_SKIP_ALL") == "1":       # dev knob: the multi-rank step without its collective
                return
            self._zero_unused()                                 # stale slots must not be summed into other ranks' gradients
            if self.dp_mode == "nvls":
                nv = self._nvls
                if nv["overlap"] and nv["launched"] > 0:
                    # join: every bucket's exchange was issued on the side stream as its gradients completed
                    done = torch.cuda.Event()
                    done.record(nv["stream"])
                    torch.cuda.current_stream().wait_event(done)
                    assert nv["launched"] == len(self.bucket_names), "a gradient bucket never completed"
                    nv["launched"] = 0
                else:
                    self._nvls_all_reduce()
                return
            buf = self.flat_g
            if self.reduce_dtype == torch.float32:
                dist.all_reduce(buf, op=dist.ReduceOp.SUM, group=self.pg)
            else:
                low = buf.to(self.reduce_dtype)
                dist.all_reduce(low, op=dist.ReduceOp.SUM, group=self.pg)
                buf.copy_(low)
            return
        for w, low, buf in self._works:
            w.wait()
            if low is not None:
                buf.copy_(low)
        self._works = []

    # ------------------------------------------------------------------ the step
    def _zero_unused(self):
        # parameters whose gradient never arrived (unused) must not contribute stale data
        for bi, left in enumerate(self._pending):
            if left > 0:
                for n in self.bucket_names[bi]:
                    if self._named[n].grad is None:
                        self.slots[n][2].zero_()

    def advance(self):
        """Host side of a step: bump the step counter and publish this step's scalars (lr and the
        Adam bias corrections) to the device buffer the graph-safe kernels read."""
        self.step_count += 1
        b1, b2 = self.betas
        # pinned staging ring: the copy is asynchronous, so a slot is rewritten only after the copy that last read it
        if self._hyper_host is None:
            self._hyper_host = torch.empty(8, 4, dtype=torch.float32).pin_memory()
            self._hyper_done = [None] * 8
        k = self.step_count % 8
        if self._hyper_done[k] is not None:
            self._hyper_done[k].synchronize()
        h = self._hyper_host[k]
        h[0], h[1], h[2], h[3] = self.lr, 1.0 - b1 ** self.step_count, math.sqrt(1.0 - b2 ** self.step_count), 0.0
        self._hyper.copy_(h, non_blocking=True)
        ev = torch.cuda.Event()
        ev.record()
        self._hyper_done[k] = ev

    def step(self, from_device_scalars: bool = False):
        """Global-norm clip + Adam over the flat buffers.  Gradients are SUMS over ranks at this
        point; the 1/world average is folded into the kernel's grad_scale.
        ``from_device_scalars``: launch only (lr / bias corrections come from the device buffer that
        ``advance()`` fills) -- the form a captured CUDA graph replays."""
        self.finish_reduce()
        self._zero_unused()
        st = torch.cuda.current_stream().cuda_stream
        n = self.numel
        if self.clip > 0:
            call("bvae_grad_sumsq", self.flat_g.data_ptr(), n, self._sumsq.data_ptr(), self._ws.data_ptr(),
                 self._ws.numel(), st)
        sumsq = self._sumsq.data_ptr() if self.clip > 0 else None
        if from_device_scalars:
            call("bvae_adam_clip_step_dev", self.flat_p.data_ptr(), self.flat_g.data_ptr(), self.flat_m.data_ptr(),
                 self.flat_v.data_ptr(), n, sumsq, 1.0 / self.world, self.clip, self._hyper.data_ptr(),
                 self.betas[0], self.betas[1], self.eps, self.weight_decay, 1 if self.decoupled_wd else 0, st)
        else:
            self.step_count += 1
            call("bvae_adam_clip_step", self.flat_p.data_ptr(), self.flat_g.data_ptr(), self.flat_m.data_ptr(),
                 self.flat_v.data_ptr(), n, sumsq, 1.0 / self.world, self.clip, self.lr, self.betas[0],
                 self.betas[1], self.eps, self.weight_decay, 1 if self.decoupled_wd else 0, self.step_count, st)
        if hasattr(self.model, "invalidate_packs"):
            self.model.invalidate_packs()                      # parameters changed through raw pointers

    def grad_norm(self) -> torch.Tensor:
        """||g|| of the averaged gradient of the last step (device scalar; reading it syncs)."""
        return self._sumsq.sqrt() / self.world

    # ------------------------------------------------------------------ torch.optim.Adam-compatible state
    def optimizer_state_dict(self):
        """Same structure torch.optim.Adam.state_dict() produces (parameters() order), so the
        reference's checkpoint code and ``optimizer.load_state_dict`` interoperate."""
        names = [n for n, _ in self.model.named_parameters()]
        state = {}
        if self.step_count > 0:
            for i, n in enumerate(names):
                o, s, _ = self.slots[n]
                shape = self._named[n].shape
                state[i] = {"step": torch.tensor(float(self.step_count)),
                            "exp_avg": self.flat_m[o:o + s].view(shape).clone(),
                            "exp_avg_sq": self.flat_v[o:o + s].view(shape).clone()}
        group = {"lr": self.lr, "betas": self.betas, "eps": self.eps, "weight_decay": self.weight_decay,
                 "amsgrad": False, "maximize": False, "foreach": None, "capturable": False,
                 "differentiable": False, "fused": None, "decoupled_weight_decay": self.decoupled_wd,
                 "params": list(range(len(names)))}
        return {"state": state, "param_groups": [group]}

    def load_optimizer_state_dict(self, sd):
        names = [n for n, _ in self.model.named_parameters()]
        st = sd.get("state", {})
        step = 0
        with torch.no_grad():
            for i, n in enumerate(names):
                ent = st.get(i, st.get(str(i)))
                if ent is None:
                    continue
                o, s, _ = self.slots[n]
                self.flat_m[o:o + s].copy_(ent["exp_avg"].reshape(-1).to(self.device))
                self.flat_v[o:o + s].copy_(ent["exp_avg_sq"].reshape(-1).to(self.device))
                step = max(step, int(float(ent["step"])))
        self.step_count = step
        groups = sd.get("param_groups", [])
        if groups:
            self.lr = float(groups[0].get("lr", self.lr))

    def close(self):
        for h in self._hooks:
            h.remove()
        ops.set_grad_arena(None)
        ops.set_gn_param_deferral(True)


class StagedBatch:
    """A batch whose host->device copy is in flight on the copy stream (``GraphedTrainStep.prefetch``)."""
    __slots__ = ("slot", "buf", "ready")

    def __init__(self, slot, buf, ready):
        self.slot, self.buf, self.ready = slot, buf, ready


class GraphedTrainStep:
    """One whole training step -- zero_grad, ``BetaVAE.loss`` under autocast, backward, bucketed
    gradient all-reduce, clip + Adam -- captured once into a CUDA graph and replayed.

    The reference's loop issues the same sequence from Python every step
    (src/training/train.py:313-333); at the configured batch of 32 the B200 finishes the kernels
    faster than Python can enqueue them, so the launch sequence is recorded once.  Inputs live
    in static device buffers (``x``, ``eps``); per-step host scalars (lr, Adam bias corrections)
    are read from a device buffer; anything baked into kernel arguments (beta, capacity,
    free_bits -- they change per epoch) triggers a re-capture when it changes.

    The first ``eager_steps`` calls run un-captured (they are real steps and warm up lazy
    initialisation: cuFFT plans, smem attributes); the next call captures, and every call from
    then on replays.  Returns ``model.loss``'s dictionary; its tensors are the graph's static
    outputs and are overwritten by the next call.

    ``prefetch(x_host)`` starts the next batch's host->device copy on a separate stream (two staging
    buffers), so the PCIe transfer overlaps the current step's kernels the way the reference's
    ``DataLoader(pin_memory=True)`` + ``non_blocking`` copy overlaps its Python-bound step; pass the
    returned ``StagedBatch`` to ``__call__``."""

    def __init__(self, engine: "TrainEngine", batch_shape, autocast_dtype=torch.bfloat16, eager_steps: int = 2,
                 seed: int = 0):
        self.engine, self.model = engine, engine.model
        dev = engine.device
        self.x = torch.zeros(batch_shape, dtype=torch.float32, device=dev)
        self.eps = torch.zeros(batch_shape[0], self.model.latent_dim, dtype=torch.float32, device=dev)
        self.autocast_dtype = autocast_dtype
        self.eager_steps = eager_steps
        self.calls = 0
        self.graph = None
        self.out = None
        self._kw = None
        self.launches_per_step = None
        self._gen = torch.Generator(device=dev).manual_seed(seed)
        self._copy_stream = None
        self._stage = None
        self._consumed = [None, None]
        self._slot = 0

    def prefetch(self, x, augment=None) -> "StagedBatch":
        """Start the host->device copy of batch ``x`` on the copy stream.  ``augment`` (bvae_rt.augment.DeviceAugment)
        is applied to uint8 batches right behind the copy, on the same stream -- the reference does this work per
        sample in its DataLoader workers (augmentations.py:5-18)."""
        if self._copy_stream is None:
            self._copy_stream = torch.cuda.Stream(device=self.engine.device)
            self._stage = {}
        if x.dtype not in self._stage:               # fp32 images, or uint8 bytes (converted on the device: x / 255)
            self._stage[x.dtype] = [torch.empty(self.x.shape, dtype=x.dtype, device=self.x.device) for _ in range(2)]
        aug = augment is not None and x.dtype == torch.uint8
        if aug and "aug" not in self._stage:
            self._stage["aug"] = [torch.empty(self.x.shape, dtype=torch.uint8, device=self.x.device) for _ in range(2)]
        slot = self._slot
        self._slot ^= 1
        buf = self._stage[x.dtype][slot]
        cs = self._copy_stream
        if self._consumed[slot] is not None:
            cs.wait_event(self._consumed[slot])      # the step that read this buffer two batches ago
        with torch.cuda.stream(cs):
            buf.copy_(x, non_blocking=True)
            if aug:
                buf = augment(buf, out=self._stage["aug"][slot])
            ready = torch.cuda.Event()
            ready.record(cs)
        return StagedBatch(slot, buf, ready)

    def _load_input(self, x):
        if isinstance(x, StagedBatch):
            cur = torch.cuda.current_stream()
            cur.wait_event(x.ready)
            if x.buf.dtype == torch.uint8:
                ops.u8_to_unit_f32(x.buf, self.x)    # bytes -> ToTensor's floats, into the graph's static input
            else:
                self.x.copy_(x.buf, non_blocking=True)   # device-to-device into the graph's static input
            done = torch.cuda.Event()
            done.record(cur)
            self._consumed[x.slot] = done
        elif x.dtype == torch.uint8:
            ops.u8_to_unit_f32(x.to(self.x.device, non_blocking=True).contiguous(), self.x)
        else:
            self.x.copy_(x, non_blocking=True)

    def _body(self, kw):
        self.engine.zero_grad()
        if self.autocast_dtype is None:
            out = self.model.loss(self.x, eps=self.eps, **kw)
        else:
            with torch.autocast("cuda", dtype=self.autocast_dtype):
                out = self.model.loss(self.x, eps=self.eps, **kw)
        out["total"].backward()
        self.engine.step(from_device_scalars=True)
        # hand back plain tensors: a live autograd graph would pin this iteration's AccumulateGrad
        # nodes (and their stream) and break the capture that follows the eager steps
        # (the scalars a training loop logs arrive packed in out["scalars"]: one 16-byte device->host copy reads them all)
        return {k: (v.detach() if isinstance(v, torch.Tensor) else v) for k, v in out.items()}

    def _capture(self, kw):
        from .lib import query
        torch.cuda.synchronize()
        g = torch.cuda.CUDAGraph()
        n0 = query("bvae_launch_count")
        self.engine.zero_grad()
        with torch.cuda.graph(g):
            out = self._body(kw)
        self.launches_per_step = query("bvae_launch_count") - n0
        self.graph, self.out, self._kw = g, out, dict(kw)

    def __call__(self, x, eps=None, **kw):
        self._load_input(x)
        if eps is None:
            self.eps.normal_(generator=self._gen)
        else:
            self.eps.copy_(eps, non_blocking=True)
        self.engine.advance()
        self.calls += 1
        if self.calls <= self.eager_steps:
            return self._body(kw)
        if self.graph is None or kw != self._kw:
            self._capture(kw)
        self.graph.replay()
        self.model.invalidate_packs()      # weight packs inside the graph pool are one step old for eager callers
        return self.out


def cosine_lr(base_lr: float, epoch: int, t_max: int, eta_min: float = 0.0) -> float:
    """Closed form of torch.optim.lr_scheduler.CosineAnnealingLR after ``epoch`` steps."""
    return eta_min + (base_lr - eta_min) * (1 + math.cos(math.pi * epoch / t_max)) / 2
