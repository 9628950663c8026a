"""development: a few K3 launches for ncu (float32 Lanczos4, 4096x3000, n frames per launch)"""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch, cv2
from imgprocessor_b200 import engine, _lib
H, W = 3000, 4096
e = engine.Engine(H, W, 0)
e.set_option(_lib.OPT_K3_VARIANT, int(os.environ.get('K3V', '0')))
quad = np.float32([[60, 40], [4040, 75], [4000, 2950], [30, 2900]])
M = cv2.getPerspectiveTransform(quad, np.float32([[0, 0], [W, 0], [W, H], [0, H]]))
n = int(sys.argv[1]) if len(sys.argv) > 1 else 4
src = torch.rand((n, H, W), dtype=torch.float32, device='cuda')
for i in range(3):
    e.warp_perspective(src, M, (W, H))
torch.cuda.synchronize()
