"""Device time of the augmentation kernel and host time of its parameter draws (dev tool)."""
import sys, time
import os
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "beta-vae_b200", "src"), os.path.join(ROOT, "beta-vae_b200")]
import torch
from bvae_rt.augment import DeviceAugment, pack_params
x = torch.randint(0, 256, (32, 1, 128, 128), dtype=torch.uint8, device="cuda")
aug = DeviceAugment(True, 10, 0.1)
out = torch.empty_like(x)
for order in (True, False):
    aug.reference_order = order
    t0 = time.perf_counter()
    for _ in range(50): p = aug.draw(32)
    t1 = time.perf_counter()
    for _ in range(50): w = pack_params(*p, w=128, h=128)
    t2 = time.perf_counter()
    print(f"reference_order={order}: draw {1e6*(t1-t0)/50:.0f} us/batch, pack {1e6*(t2-t1)/50:.0f} us/batch (host)")
words = pack_params(*p, w=128, h=128).cuda()
from bvae_rt import augment as A
for _ in range(5): A.apply(x, words, out=out)
torch.cuda.synchronize()
g = torch.cuda.CUDAGraph()
with torch.cuda.graph(g):
    for _ in range(50): A.apply(x, words, out=out)
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
g.replay(); torch.cuda.synchronize()
e0.record(); g.replay(); e1.record(); torch.cuda.synchronize()
print(f"augment_u8 kernel, batch 32 x 1 x 128 x 128: {1e3*e0.elapsed_time(e1)/50:.2f} us per launch (graph of 50)")
x2 = torch.randint(0, 256, (1024, 1, 128, 128), dtype=torch.uint8, device="cuda"); o2 = torch.empty_like(x2)
w2 = pack_params(*DeviceAugment(True,10,0.1,reference_order=False).draw(1024), w=128, h=128).cuda()
for _ in range(3): A.apply(x2, w2, out=o2)
torch.cuda.synchronize(); e0.record()
for _ in range(20): A.apply(x2, w2, out=o2)
e1.record(); torch.cuda.synchronize()
print(f"batch 1024: {1e3*e0.elapsed_time(e1)/20:.1f} us per launch = {2*x2.numel()/(e0.elapsed_time(e1)/20*1e-3)/1e9:.0f} GB/s (read + write)")
