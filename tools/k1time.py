"""development: time K1 variant 4 (stream2) with the currently loaded library"""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from imgprocessor_b200 import _lib, engine, synth
H, W = 3000, 4096
e = engine.Engine(H, W, 0)
e.set_dark(synth.dark_map(H, W)); e.set_flat(synth.flat_map(H, W))
n = 8
raw = synth.scene_torch(n, H, W, 7, torch.device('cuda', 0), 'uint16')
out = torch.empty((n, H, W), dtype=torch.float32, device='cuda')
res = []
for variant, segs in ((3, (0, 32, 64, 128)),):
    for seg in segs:
        e.set_option(_lib.OPT_K1_VARIANT, variant); e.set_option(_lib.OPT_K1_SEG_ROWS, seg)
        for i in range(3): e.pointwise_median(raw[i % n], 0.1, 3, out=out[i % n:i % n + 1])
        torch.cuda.synchronize()
        ev = [torch.cuda.Event(enable_timing=True) for _ in range(33)]
        ev[0].record()
        for i in range(32):
            e.pointwise_median(raw[i % n], 0.1, 3, out=out[i % n:i % n + 1]); ev[i + 1].record()
        torch.cuda.synchronize()
        ts = sorted(ev[i].elapsed_time(ev[i + 1]) * 1e3 for i in range(32))
        res.append('v%d seg%d: %.1f us' % (variant, seg, ts[16]))
    # 4 frames in one launch
e.set_option(_lib.OPT_K1_VARIANT, 3); e.set_option(_lib.OPT_K1_SEG_ROWS, 0)
for i in range(2): e.pointwise_median(raw[:4], 0.1, 3, out=out[:4])
torch.cuda.synchronize()
a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
a.record()
for i in range(8): e.pointwise_median(raw[4 * (i % 2):4 * (i % 2) + 4], 0.1, 3, out=out[:4])
b.record(); torch.cuda.synchronize()
res.append('v3 4-frame launch: %.1f us/frame' % (a.elapsed_time(b) * 1e3 / 32))
print(os.environ.get('IMGCORR_LIB', 'default'), ' | '.join(res))
