"""Host-side per-epoch scalars: KL weight (beta) and KL capacity target.

Drop-in for the reference's ``training.schedulers`` (reference: src/training/schedulers.py:23-108):
same class names, constructor arguments and ``value(epoch)`` results.  The values are kernel
arguments of ``bvae_heads_reparam_kl_fwd`` (beta, capacity); a change re-captures the step graph."""
import math


def _section(cfg, name, default=None):
    """Child ``name`` of a dict / attribute-style config as a plain dict (``default`` when absent)."""
    if cfg is None:
        return default
    child = cfg.get(name, None) if isinstance(cfg, dict) else getattr(cfg, name, None)
    if child is None:
        return default
    if isinstance(child, dict):
        return child
    if hasattr(child, "to_dict"):
        return child.to_dict()
    return {k: v for k, v in vars(child).items() if not k.startswith("_")}


def _first(mapping, *keys, default=None):
    for k in keys:
        if k in mapping and mapping[k] is not None:
            return mapping[k]
    return default


class BetaScheduler:
    """``beta_schedule``: type constant | linear | cosine | cyclical with start/end beta, warm-up and cycle
    length (long and short key spellings); without the section the constant ``model.beta``."""

    def __init__(self, root_cfg, total_epochs):
        sched = _section(root_cfg, "beta_schedule")
        if sched is None:
            model = _section(root_cfg, "model", {}) or {}
            sched = {"type": "constant", "end_beta": model.get("beta", 1.0)}
        self.type = sched.get("type", "constant")
        end_guess = _first(sched, "end_beta", "end", default=1.0)
        self.start = _first(sched, "start_beta", "start", default=end_guess)
        self.end = _first(sched, "end_beta", "end", default=self.start)
        self.warm = _first(sched, "warmup_epochs", "warmup", default=0)
        self.cycle = _first(sched, "cycle_length", "cycle", default=0)
        self.total_epochs = total_epochs

    def value(self, epoch: int):
        span = self.end - self.start
        if self.type == "linear" and self.warm > 0:
            return self.start + span * min(1.0, epoch / float(self.warm))
        if self.type == "cosine" and self.total_epochs > 1:
            return self.start + 0.5 * span * (1.0 - math.cos(math.pi * epoch / (self.total_epochs - 1)))
        if self.type in ("cyclical", "cyc") and self.cycle > 0:
            return self.start + span * ((epoch % self.cycle) / float(self.cycle))
        return self.end


class CapacityScheduler:
    """``loss.capacity_schedule``: linear ramp C_start -> C_end over ``warmup_epochs`` then flat; ``None``
    while disabled (the loss then runs in beta mode)."""

    def __init__(self, root_cfg, total_epochs):
        loss = _section(root_cfg, "loss", {}) or {}
        cs = loss.get("capacity_schedule", {}) or {}
        self.enabled = bool(cs.get("enabled", False))
        self.C0 = cs.get("C_start", 0.0)
        self.C1 = cs.get("C_end", self.C0)
        self.warm = cs.get("warmup_epochs", 0)
        self.total = cs.get("total_epochs", total_epochs)
        self.total_epochs = total_epochs

    def value(self, epoch: int):
        if not self.enabled:
            return None
        e = max(0, epoch)
        if e > self.warm:
            return self.C1
        return self.C0 + min(1.0, e / max(1, self.warm)) * (self.C1 - self.C0)
