"""development: run one K1 variant a few times (for ncu)"""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from imgprocessor_b200 import _lib, engine, synth
variant, seg = int(sys.argv[1]), int(sys.argv[2])
H, W = 3000, 4096
e = engine.Engine(H, W, 0)
e.set_dark(synth.dark_map(H, W)); e.set_flat(synth.flat_map(H, W))
raw = synth.scene_torch(4, H, W, 7, torch.device('cuda', 0), 'uint16')
out = torch.empty((1, H, W), dtype=torch.float32, device='cuda')
e.set_option(_lib.OPT_K1_VARIANT, variant); e.set_option(_lib.OPT_K1_SEG_ROWS, seg)
for i in range(6):
    e.pointwise_median(raw[i % 4], 0.1, int(os.environ.get("KSIZE", "3")), out=out)
torch.cuda.synchronize()
