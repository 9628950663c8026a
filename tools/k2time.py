"""development: K2 (lens undistortion) timings with the currently loaded library (IMGCORR_LIB selects a variant build):
one frame per launch, 8 and 32 frames per launch, for a lens preset / frame shape.
usage: python tools/k2time.py [moderate|strong] [H W]"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import cv2
import numpy as np
import torch
from imgprocessor_b200 import _lib, engine, synth

lens = sys.argv[1] if len(sys.argv) > 1 else 'moderate'
H, W = (int(sys.argv[2]), int(sys.argv[3])) if len(sys.argv) > 3 else (3000, 4096)
e = engine.Engine(H, W, 0)
p = synth.lens_moderate(H, W) if lens == 'moderate' else synth.lens_strong(H, W)
K, d = synth.camera_matrix(p), synth.dist_coeffs(p)
P, roi = cv2.getOptimalNewCameraMatrix(K, d, (W, H), 1, (W, H))
e.set_lens(K, d, P)
n = 32 if H * W <= 3000 * 4096 else 4
src = torch.rand((n, H, W), dtype=torch.float32, device='cuda')
out = torch.empty((n, H, W), dtype=torch.float32, device='cuda')
res = []
kv = int(os.environ.get('K2V', '0'))
e.set_option(_lib.OPT_K2_VARIANT, kv)


def t(fn, iters=24):
    for i in range(3):
        fn(i)
    torch.cuda.synchronize()
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(iters + 1)]
    torch.cuda._sleep(40000000)          # ~20 ms: the launches below queue up behind it, so the events time the GPU, not Python
    ev[0].record()
    for i in range(iters):
        fn(i)
        ev[i + 1].record()
    torch.cuda.synchronize()
    return sorted(ev[i].elapsed_time(ev[i + 1]) * 1e3 for i in range(iters))[iters // 2]


res.append('1 frame: %.1f us' % t(lambda i: e.undistort(src[i % n], out=out[i % n:i % n + 1])))
if n >= 8:
    res.append('8/launch: %.1f us/frame' % (t(lambda i: e.undistort(src[8 * (i % (n // 8)):8 * (i % (n // 8)) + 8], out=out[:8]), 8) / 8))
res.append('%d/launch: %.1f us/frame' % (n, t(lambda i: e.undistort(src, out=out), 6) / n))
x0, y0, w, h = (int(v) for v in roi)
res.append('roi crop 1 frame: %.1f us' % t(lambda i: e.undistort(src[i % n], window=(x0, y0, w, h))))
print('k2 variant', kv, lens, '%dx%d' % (H, W), os.environ.get('IMGCORR_LIB', 'default').split('/')[-1], ' | '.join(res), flush=True)
