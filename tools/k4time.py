"""development: time K4 (STE-free average) with the currently loaded library"""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from imgprocessor_b200 import engine
H, W = 3000, 4096
e = engine.Engine(H, W, 0)
res = []
for n in (2, 4, 8):
    base = torch.rand((H, W), device='cuda') * 3000 + 500
    fr = (base[None] + torch.randn((n, H, W), device='cuda') * 0.5 * base.sqrt()[None]
          + (torch.rand((n, H, W), device='cuda') < 1e-3) * 3000).clamp(0, 65535).to(torch.uint16)
    for i in range(2): e.ste_average(fr, (5.0, 0.0, 0.5), 4)
    torch.cuda.synchronize()
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(9)]
    ev[0].record()
    for i in range(8):
        e.ste_average(fr, (5.0, 0.0, 0.5), 4); ev[i + 1].record()
    torch.cuda.synchronize()
    t = sorted(ev[i].elapsed_time(ev[i + 1]) * 1e3 for i in range(8))[4]
    res.append('%d exposures: %.0f us (%.0f us per launch)' % (n, t, t / (n - 1)))
print('k4 uint16 4096x3000 |', ' | '.join(res))
