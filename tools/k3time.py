"""development: time K3 (perspective warp) with the currently loaded library"""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch, cv2
from imgprocessor_b200 import _lib, engine
H, W = 3000, 4096
e = engine.Engine(H, W, 0)
quad = np.float32([[60, 40], [4040, 75], [4000, 2950], [30, 2900]])
M = cv2.getPerspectiveTransform(quad, np.float32([[0, 0], [W, 0], [W, H], [0, H]]))
n = 8
def t(fn, iters=12):
    for i in range(2): fn(i)
    torch.cuda.synchronize()
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(iters + 1)]
    ev[0].record()
    for i in range(iters):
        fn(i); ev[i + 1].record()
    torch.cuda.synchronize()
    return sorted(ev[i].elapsed_time(ev[i + 1]) * 1e3 for i in range(iters))[iters // 2]
res = []
for dt in (torch.float32, torch.uint16, torch.float64):
    src = (torch.rand((n, H, W), device='cuda') * 60000).to(dt)
    for interp in ('lanczos4', 'cubic'):
        a = t(lambda i: e.warp_perspective(src[i % n], M, (W, H), interp))
        b = t(lambda i: e.warp_perspective(src, M, (W, H), interp), 4) / n
        res.append('%s %s: %.0f us single, %.0f us/frame x%d' % (str(dt).split('.')[-1], interp, a, b, n))
    del src
print('k3', os.environ.get('IMGCORR_LIB', 'default').split('/')[-1], ' | '.join(res))
