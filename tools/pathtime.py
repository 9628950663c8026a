"""development: timings of the secondary paths (direct medianThreshold / LensDistortion.correct calls on 4096x3000 frames)"""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch, cv2
from imgprocessor_b200 import _lib, engine, synth
H, W, n = 3000, 4096, 8
e = engine.Engine(H, W, 0)
p = synth.lens_moderate(H, W)
K, d = synth.camera_matrix(p), synth.dist_coeffs(p)
P, roi = cv2.getOptimalNewCameraMatrix(K, d, (W, H), 1, (W, H))
e.set_lens(K, d, P)
def t(fn, iters=10):
    for i in range(3): fn(i)
    torch.cuda.synchronize()
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(iters + 1)]
    ev[0].record()
    for i in range(iters):
        fn(i); ev[i + 1].record()
    torch.cuda.synchronize()
    return sorted(ev[i].elapsed_time(ev[i + 1]) * 1e3 for i in range(iters))[iters // 2]
res = []
for dt in (torch.uint16, torch.float32, torch.uint8):
    img = (torch.rand((n, H, W), device='cuda') * (250 if dt == torch.uint8 else 60000)).to(dt)
    out = torch.empty_like(img)
    name = str(dt).split('.')[-1]
    for ks in (3, 5):
        a = t(lambda i: e.pointwise_median(img[i % n], 0.1, ks, flags=0, out_dtype=dt, want_mask=True))
        b = t(lambda i: e.pointwise_median(img, 0.1, ks, flags=0, out_dtype=dt, want_mask=True), 4) / n
        res.append('medianThreshold %s %dx%d: %.0f us single, %.0f us/frame x%d' % (name, ks, ks, a, b, n))
    a = t(lambda i: e.undistort(img[i % n]))
    b = t(lambda i: e.undistort(img), 4) / n
    res.append('LensDistortion.correct %s: %.0f us single, %.0f us/frame x%d' % (name, a, b, n))
    del img, out
print('\n'.join(res))
