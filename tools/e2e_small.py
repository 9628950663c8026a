"""development: end-to-end (host-buffer) throughput for small frames (BASELINE configs[0]: 1024x1024 float32)"""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch, cv2
from imgprocessor_b200 import engine, synth
for (H, W, dt) in ((1024, 1024, np.float32), (512, 640, np.uint16)):
    e = engine.Engine(H, W, 0)
    e.set_dark(synth.dark_map(H, W)); e.set_flat(synth.flat_map(H, W))
    p = synth.lens_moderate(H, W)
    K, d = synth.camera_matrix(p), synth.dist_coeffs(p)
    P, roi = cv2.getOptimalNewCameraMatrix(K, d, (W, H), 1, (W, H))
    e.set_lens(K, d, P)
    n = 512
    h_in = engine.pinned_empty((n, H, W), dt)
    h_out = engine.pinned_empty((n, H, W), np.float32)
    h_in[...] = synth.scene(H, W, 3, dt)[None]
    e.correct_host(h_in, out=h_out)
    t0 = time.perf_counter()
    for r in range(3): e.correct_host(h_in, out=h_out)
    dt_s = (time.perf_counter() - t0) / 3
    bytes_per_frame = H * W * (np.dtype(dt).itemsize + 4)
    print('%dx%d %s: %.1f us per frame end to end = %.0f Mpx/s (PCIe time at 50 GB/s D2H: %.1f us)' % (H, W, np.dtype(dt).name, dt_s / n * 1e6, n * H * W / dt_s / 1e6, H * W * 4 / 50e3))
