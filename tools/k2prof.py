import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch, cv2
from imgprocessor_b200 import engine, synth
H, W = 3000, 4096
e = engine.Engine(H, W, 0)
p = synth.lens_moderate(H, W)
K, d = synth.camera_matrix(p), synth.dist_coeffs(p)
P, roi = cv2.getOptimalNewCameraMatrix(K, d, (W, H), 1, (W, H))
e.set_lens(K, d, P)
n = int(sys.argv[1]) if len(sys.argv) > 1 else 4
src = torch.rand((n, H, W), dtype=torch.float32, device='cuda')
out = torch.empty((n, H, W), dtype=torch.float32, device='cuda')
for i in range(3):
    e.undistort(src, out=out)
torch.cuda.synchronize()
