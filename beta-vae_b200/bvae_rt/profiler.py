"""Per-op CUDA-event timing on the launching stream (bench.py's roofline numbers).

Disabled by default (a no-op context); when enabled each C-ABI call that ``ops`` makes is
bracketed by two events recorded on torch's current stream -- the stream the kernels are
launched on -- together with the algorithmic FLOPs / bytes the caller states for it."""
import contextlib

import torch

from . import lib

_active = None


class OpProfile:
    def __init__(self):
        self.records = []  # (tag, start_event, end_event, flops, bytes, launches)

    def summary(self):
        torch.cuda.synchronize()
        agg = {}
        for tag, e0, e1, flops, nbytes, launches in self.records:
            a = agg.setdefault(tag, {"ms": 0.0, "calls": 0, "flops": 0.0, "bytes": 0.0, "launches": 0})
            a["ms"] += e0.elapsed_time(e1)
            a["calls"] += 1
            a["flops"] += flops
            a["bytes"] += nbytes
            a["launches"] += launches
        return agg


def start():
    global _active
    _active = OpProfile()
    return _active


def stop():
    global _active
    p, _active = _active, None
    return p


@contextlib.contextmanager
def region(tag, flops=0.0, nbytes=0.0):
    if _active is None:
        yield
        return
    e0 = torch.cuda.Event(enable_timing=True)
    e1 = torch.cuda.Event(enable_timing=True)
    n0 = lib.query("bvae_launch_count")
    e0.record()
    try:
        yield
    finally:
        e1.record()
        _active.records.append((tag, e0, e1, float(flops), float(nbytes), lib.query("bvae_launch_count") - n0))
