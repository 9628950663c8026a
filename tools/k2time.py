"""development: time K2 with the currently loaded library"""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch, cv2
from imgprocessor_b200 import _lib, engine, synth
H, W = 3000, 4096
e = engine.Engine(H, W, 0)
p = synth.lens_moderate(H, W)
K, d = synth.camera_matrix(p), synth.dist_coeffs(p)
P, roi = cv2.getOptimalNewCameraMatrix(K, d, (W, H), 1, (W, H))
e.set_lens(K, d, P)
n = 8
src = torch.rand((n, H, W), dtype=torch.float32, device='cuda')
out = torch.empty((n, H, W), dtype=torch.float32, device='cuda')
res = []
kv = int(os.environ.get('K2V', '0'))
e.set_option(_lib.OPT_K2_VARIANT, kv)
def t(fn, iters=24):
    for i in range(3): fn(i)
    torch.cuda.synchronize()
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(iters + 1)]
    ev[0].record()
    for i in range(iters):
        fn(i); ev[i + 1].record()
    torch.cuda.synchronize()
    return sorted(ev[i].elapsed_time(ev[i + 1]) * 1e3 for i in range(iters))[iters // 2]
res.append('1 frame: %.1f us' % t(lambda i: e.undistort(src[i % n], out=out[i % n:i % n + 1])))
res.append('4 frames/launch: %.1f us/frame' % (t(lambda i: e.undistort(src[4 * (i % 2):4 * (i % 2) + 4], out=out[:4]), 8) / 4))
res.append('8 frames/launch: %.1f us/frame' % (t(lambda i: e.undistort(src, out=out), 6) / 8))
mx, my = e.undistort_maps()
res.append('explicit maps: %.1f us' % t(lambda i: e.remap(src[i % n], mx, my)))
print('k2 variant', kv, os.environ.get('IMGCORR_LIB', 'default').split('/')[-1], ' | '.join(res))
