"""development: time the pointwise-only K1 path (threshold <= 0: dark + flat only) and the K1 mask / direct variants"""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from imgprocessor_b200 import _lib, engine, synth
H, W, n = 3000, 4096, 16
e = engine.Engine(H, W, 0)
e.set_dark(synth.dark_map(H, W)); e.set_flat(synth.flat_map(H, W))
raw = synth.scene_torch(n, H, W, 7, torch.device('cuda', 0), 'uint16')
out = torch.empty((n, H, W), dtype=torch.float32, device='cuda')
def t(fn, iters=10):
    for i in range(3): fn(i)
    torch.cuda.synchronize()
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(iters + 1)]
    ev[0].record()
    for i in range(iters):
        fn(i); ev[i + 1].record()
    torch.cuda.synchronize()
    return sorted(ev[i].elapsed_time(ev[i + 1]) * 1e3 for i in range(iters))[iters // 2]
a = t(lambda i: e.pointwise_median(raw[i % n], 0.0, 0, out=out[i % n:i % n + 1]))
b = t(lambda i: e.pointwise_median(raw, 0.0, 0, out=out), 6) / n
c = t(lambda i: e.pointwise_median(raw, 0.1, 3, out=out), 6) / n
print('pointwise only: %.1f us single, %.1f us/frame x%d (14 B/px -> %.0f GB/s) | with 3x3 median: %.1f us/frame' % (a, b, n, 14 * H * W / b / 1e3, c))
