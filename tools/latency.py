"""Latency of one drop-in call: CameraCalibration.correct(frame) on one 4096x3000 uint16 frame (float64 result), full chain.
usage: python tools/latency.py [H W]   -> median / min ms over 20 calls, with writable maps (fingerprinted every call) and
read-only maps (keyed by identity)."""
import sys
import time
import os
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from imgprocessor_b200.camera import CameraCalibration      # noqa: E402


def main():
    H, W = (int(sys.argv[1]), int(sys.argv[2])) if len(sys.argv) > 2 else (3000, 4096)
    rng = np.random.default_rng(0)
    dark = rng.normal(100, 3, (H, W))
    flat = np.clip(rng.normal(0.9, 0.05, (H, W)), 0.3, 1.2)
    frame = rng.integers(90, 4000, (H, W)).astype(np.uint16)
    for ro in (False, True):
        cal = CameraCalibration()
        cal.addDarkCurrent(dark.copy())
        cal.addFlatField(flat.copy())
        cal.addLens(_lens(H, W))
        if ro:
            for k in ('dark current', 'flat field'):
                for entry in _entries(cal.coeffs[k]):
                    for a in entry:
                        if isinstance(a, np.ndarray):
                            a.setflags(write=False)
        ts = []
        import io
        import contextlib
        for i in range(24):
            t0 = time.perf_counter()
            with contextlib.redirect_stdout(io.StringIO()):
                cal.correct(frame, threshold=0.1)
            ts.append((time.perf_counter() - t0) * 1e3)
        ts = ts[4:]
        print('%dx%d %s maps: median %.2f ms  min %.2f ms' % (W, H, 'read-only' if ro else 'writable', float(np.median(ts)), min(ts)))


def _entries(c):
    if isinstance(c, dict):
        for v in c.values():
            for e in _entries(v):
                yield e
    elif isinstance(c, (list, tuple)):
        if any(isinstance(a, np.ndarray) for a in c):
            yield c
        else:
            for v in c:
                for e in _entries(v):
                    yield e


def _lens(H, W):
    from imgprocessor_b200.camera import LensDistortion
    ld = LensDistortion({})
    f = 1.2 * W
    ld.setCameraParams(f, f, W / 2.0, H / 2.0, -0.12, 0.05, 0.01, 0.0005, -0.0004)
    ld._coeffs['shape'] = (H, W)
    return ld


if __name__ == '__main__':
    main()
