"""Kernel micro-benchmark (development tool): CUDA-event timings of K1 variants, K2 and the chain on
BASELINE-sized frames, with algorithmic GB/s (K1 14 B/px u16 / 16 B/px f32, K2 8 B/px).
Rotates over several frames so inputs exceed L2.  Usage: python tools/kbench.py [--H 3000 --W 4096]"""
import argparse
import json
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from imgprocessor_b200 import _lib, engine, synth  # noqa: E402


def timeit(fn, iters, warm=3):
    for _ in range(warm):
        fn(0)
    torch.cuda.synchronize()
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(iters + 1)]
    ev[0].record()
    for i in range(iters):
        fn(i)
        ev[i + 1].record()
    torch.cuda.synchronize()
    ts = [ev[i].elapsed_time(ev[i + 1]) * 1e3 for i in range(iters)]
    return float(np.median(ts)), float(np.min(ts))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--H', type=int, default=3000)
    ap.add_argument('--W', type=int, default=4096)
    ap.add_argument('--frames', type=int, default=16)
    ap.add_argument('--iters', type=int, default=48)
    ap.add_argument('--dtype', default='uint16')
    ap.add_argument('--ksize', type=int, default=3)
    ap.add_argument('--lens', default='moderate')
    a = ap.parse_args()
    H, W, n = a.H, a.W, a.frames
    dev = torch.device('cuda', 0)
    e = engine.Engine(H, W, 0)
    e.set_dark(synth.dark_map(H, W))
    e.set_flat(synth.flat_map(H, W))
    p = synth.lens_moderate(H, W) if a.lens == 'moderate' else synth.lens_strong(H, W)
    import cv2
    K, d = synth.camera_matrix(p), synth.dist_coeffs(p)
    P, roi = cv2.getOptimalNewCameraMatrix(K, d, (W, H), 1, (W, H))
    e.set_lens(K, d, P)
    raw = synth.scene_torch(n, H, W, 7, dev, a.dtype)
    mid = torch.empty((n, H, W), dtype=torch.float32, device=dev)
    out = torch.empty((n, H, W), dtype=torch.float32, device=dev)
    px = H * W
    rb = 2 if a.dtype == 'uint16' else 4
    res = {'H': H, 'W': W, 'dtype': a.dtype, 'ksize': a.ksize}

    def rep(name, med, best, bytes_px):
        res[name] = {'us_median': round(med, 2), 'us_best': round(best, 2),
                     'GBps_median': round(bytes_px * px / med * 1e-3, 1), 'Gpx_s': round(px / med * 1e-3, 2)}
        print(name, res[name], flush=True)

    for variant, nm in ((1, 'k1_generic'), (2, 'k1_tma'), (3, 'k1_stream')):
        e.set_option(_lib.OPT_K1_VARIANT, variant)
        try:
            med, best = timeit(lambda i: e.pointwise_median(raw[i % n], 0.1, a.ksize, out=mid[i % n:i % n + 1]), a.iters)
            rep(nm, med, best, rb + 12)
        except Exception as ex:
            print(nm, 'failed:', ex)
    e.set_option(_lib.OPT_K1_SEG_ROWS, 0)
    e.set_option(_lib.OPT_K1_VARIANT, 0)
    med, best = timeit(lambda i: e.pointwise_median(raw[i % n], 0.0, 0, flags=3, out=mid[i % n:i % n + 1]), a.iters)
    rep('k1_pointwise_only', med, best, rb + 12)
    med, best = timeit(lambda i: e.undistort(mid[i % n], out=out[i % n:i % n + 1]), a.iters)
    rep('k2_undistort', med, best, 8)
    med, best = timeit(lambda i: e.undistort(mid[i % n], out_dtype=torch.float32, window=tuple(int(v) for v in roi)), a.iters)
    rep('k2_undistort_roi', med, best, 8)
    med, best = timeit(lambda i: e.correct_batch(raw[i % n], 0.1, a.ksize, out=out[i % n:i % n + 1]), a.iters)
    rep('chain_per_frame', med, best, rb + 20)
    # whole batch in one call
    med, best = timeit(lambda i: e.correct_batch(raw, 0.1, a.ksize, out=out), 6, warm=2)
    rep('chain_batch%d' % n, med / n, best / n, rb + 20)
    # plain device copy as the HBM yardstick
    big = torch.empty(256 * 1024 * 1024, dtype=torch.float32, device=dev)
    big2 = torch.empty_like(big)
    med, best = timeit(lambda i: big2.copy_(big), 10)
    res['copy_GBps'] = round(2 * big.numel() * 4 / best * 1e-3, 1)
    print('copy_GBps', res['copy_GBps'])
    os.makedirs('gpurun_out', exist_ok=True)
    with open('gpurun_out/kbench_%s_%dx%d_k%d.json' % (a.dtype, H, W, a.ksize), 'w') as f:
        json.dump(res, f, indent=1)


if __name__ == '__main__':
    main()
