"""BASELINE configs[4] (C5): 6000x4000 uint16 frames streamed from a pool of 16 pinned host buffers through the full chain
(imgcorr_correct_host); prints throughput and checks one frame against the device-resident path."""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch, cv2
from imgprocessor_b200 import engine, synth
H, W, pool, total = 4000, 6000, 16, int(sys.argv[1]) if len(sys.argv) > 1 else 256
e = engine.Engine(H, W, 0)
e.set_dark(synth.dark_map(H, W)); e.set_flat(synth.flat_map(H, W))
p = synth.lens_moderate(H, W)
K, d = synth.camera_matrix(p), synth.dist_coeffs(p)
P, roi = cv2.getOptimalNewCameraMatrix(K, d, (W, H), 1, (W, H))
e.set_lens(K, d, P)
h_in = engine.pinned_empty((pool, H, W), np.uint16)
h_out = engine.pinned_empty((pool, H, W), np.float32)
src = synth.scene_torch(pool, H, W, 3, torch.device('cuda', 0), 'uint16')
h_in[...] = src.cpu().numpy()
e.correct_host(h_in, out=h_out)
t0 = time.perf_counter()
for r in range(total // pool):
    e.correct_host(h_in, out=h_out)
dt = time.perf_counter() - t0
ref = e.correct_batch(src[:2]).cpu().numpy()
print('C5: %d frames of %dx%d in %.2f s = %.0f Mpx/s end to end; first frames identical to the device path: %s'
      % (total, W, H, dt, total * H * W / dt / 1e6, np.array_equal(ref, h_out[:2])))
t = torch.cuda.Event(enable_timing=True); t2 = torch.cuda.Event(enable_timing=True)
out = torch.empty((pool, H, W), dtype=torch.float32, device='cuda')
e.correct_batch(src, out=out); t.record()
for r in range(4): e.correct_batch(src, out=out)
t2.record(); torch.cuda.synchronize()
print('C5 device-resident: %.0f Mpx/s' % (4 * pool * H * W / (t.elapsed_time(t2) * 1e-3) / 1e6))
