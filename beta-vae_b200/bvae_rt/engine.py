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
    b3 = take("fc_logvar.") + take("fc_mu.")
    b4 = []
    for i in reversed(range(nb)):
        b4 += take(f"encoder.{i}.")
    got = b1 + b2 + b3 + b4
    assert sorted(got) == sorted(names), "bucket plan does not cover the parameter list"
    return [b1, b2, b3, b4]


class TrainEngine:
    def __init__(self, model, lr: float, betas=(0.9, 0.999), eps: float = 1e-8, weight_decay: float = 0.0,
                 decoupled_wd: bool = False, clip: float = 0.0, process_group=None, reduce_dtype=torch.float32):
        self.model = model
        self.lr, self.betas, self.eps = float(lr), tuple(betas), float(eps)
        self.weight_decay, self.decoupled_wd, self.clip = float(weight_decay), bool(decoupled_wd), float(clip)
        self.step_count = 0
        self.pg = process_group
        self.world = dist.get_world_size(process_group) if (dist.is_available() and dist.is_initialized()) else 1
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
        self.numel = off
        self.flat_p = torch.zeros(off, dtype=torch.float32, device=dev)
        self.flat_g = torch.zeros(off, dtype=torch.float32, device=dev)
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
        self._ws = torch.empty(query("bvae_adam_workspace", self.numel), dtype=torch.uint8, device=dev)
        self._hooks = []
        arena = ops.GradArena()
        for n, p in named.items():
            arena.register(p, self.slots[n][2])
            self._hooks.append(p.register_post_accumulate_grad_hook(self._make_hook(n)))
        self.arena = arena
        ops.set_grad_arena(arena)
        if hasattr(model, "invalidate_packs"):
            model.invalidate_packs()
        self.last_grad_norm = None

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
                self._launch_reduce(bi)
        return hook

    def _launch_reduce(self, bi):
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

    def finish_reduce(self):
        """Wait for the bucket all-reduces (stream-side) and restore fp32 gradients."""
        for w, low, buf in self._works:
            w.wait()
            if low is not None:
                buf.copy_(low)
        self._works = []

    # ------------------------------------------------------------------ the step
    def step(self):
        """Global-norm clip + Adam over the flat buffers.  Gradients are SUMS over ranks at this
        point; the 1/world average is folded into the kernel's grad_scale."""
        self.finish_reduce()
        # parameters whose gradient never arrived (unused) must not contribute stale data
        for bi, left in enumerate(self._pending):
            if left > 0:
                for n in self.bucket_names[bi]:
                    if self._named[n].grad is None:
                        self.slots[n][2].zero_()
        self.step_count += 1
        st = torch.cuda.current_stream().cuda_stream
        n = self.numel
        if self.clip > 0:
            call("bvae_grad_sumsq", self.flat_g.data_ptr(), n, self._sumsq.data_ptr(), self._ws.data_ptr(),
                 self._ws.numel(), st)
        call("bvae_adam_clip_step", self.flat_p.data_ptr(), self.flat_g.data_ptr(), self.flat_m.data_ptr(),
             self.flat_v.data_ptr(), n, self._sumsq.data_ptr() if self.clip > 0 else None, 1.0 / self.world,
             self.clip, self.lr, self.betas[0], self.betas[1], self.eps, self.weight_decay,
             1 if self.decoupled_wd else 0, self.step_count, st)
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


def cosine_lr(base_lr: float, epoch: int, t_max: int, eta_min: float = 0.0) -> float:
    """Closed form of torch.optim.lr_scheduler.CosineAnnealingLR after ``epoch`` steps."""
    return eta_min + (base_lr - eta_min) * (1 + math.cos(math.pi * epoch / t_max)) / 2
