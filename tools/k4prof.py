import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from imgprocessor_b200 import engine
H, W = 3000, 4096
e = engine.Engine(H, W, 0)
fr = (torch.rand((4, H, W), device='cuda') * 3000 + 500).to(torch.uint16)
for i in range(2): e.ste_average(fr, (5.0, 0.0, 0.5), 4)
torch.cuda.synchronize()
