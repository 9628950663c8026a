"""development: time the 5x5 K1 kernels with the currently loaded library (4096x3000 uint16 chain, 8192x8192 float32 direct)"""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from imgprocessor_b200 import _lib, engine, synth
def t(fn, iters=16):
    for i in range(3): fn(i)
    torch.cuda.synchronize()
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(iters + 1)]
    ev[0].record()
    for i in range(iters):
        fn(i); ev[i + 1].record()
    torch.cuda.synchronize()
    return sorted(ev[i].elapsed_time(ev[i + 1]) * 1e3 for i in range(iters))[iters // 2]
res = []
H, W = 3000, 4096
e = engine.Engine(H, W, 0)
e.set_dark(synth.dark_map(H, W)); e.set_flat(synth.flat_map(H, W))
n = 8
raw = synth.scene_torch(n, H, W, 7, torch.device('cuda', 0), 'uint16')
out = torch.empty((n, H, W), dtype=torch.float32, device='cuda')
for variant in (2, 3):
    e.set_option(_lib.OPT_K1_VARIANT, variant)
    res.append('u16 chain v%d: %.1f us single, %.1f us/frame x8' % (
        variant, t(lambda i: e.pointwise_median(raw[i % n], 0.1, 5, out=out[i % n:i % n + 1])),
        t(lambda i: e.pointwise_median(raw, 0.1, 5, out=out), 4) / n))
del raw, out, e
H = W = 8192
e = engine.Engine(H, W, 0)
img = torch.rand((H, W), device='cuda') * 4000
out = torch.empty((1, H, W), dtype=torch.float32, device='cuda')
for variant in (2, 3):
    e.set_option(_lib.OPT_K1_VARIANT, variant)
    res.append('8192^2 f32 direct v%d: %.0f us, with mask %.0f us' % (variant, t(lambda i: e.pointwise_median(img, 0.1, 5, flags=0, out=out), 8),
               t(lambda i: e.pointwise_median(img, 0.1, 5, flags=0, out=out, want_mask=True), 8)))
print(os.environ.get('IMGCORR_LIB', 'default').split('/')[-1], ' | '.join(res))
