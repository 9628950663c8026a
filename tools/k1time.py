"""development: K1 (3x3 chain configuration) timings with the currently loaded library (IMGCORR_LIB=variants/x.so selects
a variant build): one frame per launch, and 4 / 32 frames per launch (us per frame, median over launches)."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
from imgprocessor_b200 import _lib, engine, synth

H, W = 3000, 4096
dtype = sys.argv[1] if len(sys.argv) > 1 else 'uint16'
ksize = int(sys.argv[2]) if len(sys.argv) > 2 else 3
e = engine.Engine(H, W, 0)
e.set_dark(synth.dark_map(H, W))
e.set_flat(synth.flat_map(H, W))
n = 64
raw = synth.scene_torch(n, H, W, 7, torch.device('cuda', 0), dtype)
out = torch.empty((n, H, W), dtype=torch.float32, device='cuda')
res = []
for per in (1, 4, 32):
    reps = 24 if per < 32 else 10
    groups = n // per
    for i in range(3):
        e.pointwise_median(raw[:per], 0.1, ksize, out=out[:per])
    torch.cuda.synchronize()
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(reps + 1)]
    torch.cuda._sleep(40000000)          # ~20 ms: the launches below queue up behind it, so the events time the GPU, not Python
    ev[0].record()
    for i in range(reps):
        g = i % groups
        e.pointwise_median(raw[g * per:(g + 1) * per], 0.1, ksize, out=out[g * per:(g + 1) * per])
        ev[i + 1].record()
    torch.cuda.synchronize()
    ts = sorted(ev[i].elapsed_time(ev[i + 1]) * 1e3 / per for i in range(reps))
    res.append('%d/launch: %.2f us/frame (min %.2f)' % (per, ts[len(ts) // 2], ts[0]))
print(os.environ.get('IMGCORR_LIB', 'default'), dtype, 'ksize', ksize, ' | '.join(res), flush=True)
