"""Row-band evaluation of the float32 oracle chain: the output rows [r0, r1) of a full frame computed from only the
source rows they depend on, so that full-size frames (4096x3000, 6000x4000, 8192x8192) can be checked in seconds and
in bounded memory.

TEST INFRASTRUCTURE ONLY (see oracle/__init__.py): used by tests/ and by bench.py's in-run parity check.
Follows camera/CameraCalibration.py:351-459 like oracle/models.correct_chain_f32 (whose results it reproduces row for
row: tests/test_oracle_models.py::test_band_chain_equals_full_chain).
"""
import numpy as np

from . import models


def k1_band(raw, dark, flat, r0, r1, threshold=0.1, size=3):
    """rows [r0, r1) of: float32 pointwise (dark, flat, nan_to_num) + size x size median-threshold.  raw / dark / flat are
    full frames (dark / flat may be None).  Interior band edges get size//2 halo rows; frame edges use scipy's reflect."""
    H = raw.shape[0]
    do_med = threshold is not None and threshold > 0
    h = size // 2 if do_med else 0
    b0, b1 = max(r0 - h, 0), min(r1 + h, H)
    x = models.pointwise_model(raw[b0:b1], None if dark is None else dark[b0:b1], None if flat is None else flat[b0:b1],
                               nan_to_num=do_med)
    if do_med:
        x, _ = models.median_threshold_model(x, threshold, size)
    return x[r0 - b0:r1 - b0]


def chain_band(raw, dark, flat, mapx, mapy, r0, r1, threshold=0.1, size=3, border_value=0.0):
    """rows [r0, r1) of models.correct_chain_f32(raw, dark, flat, threshold, size, mapxy=(mapx, mapy)): the K1 band that
    the maps of those output rows reach into, then OpenCV's fixed-point bilinear remap from that band."""
    H = raw.shape[0]
    mx, my = mapx[r0:r1], mapy[r0:r1]
    _, iy, _, _ = models.fixed_point_coords(mx, my)
    lo, hi = int(max(iy.min(), 0)), int(min(iy.max() + 1, H - 1))
    if hi < lo:                                  # every window lies outside the frame
        lo, hi = 0, 0
    x = k1_band(raw, dark, flat, lo, hi + 1, threshold, size)
    return models.remap_model(x, mx, my, border_value, src_row0=lo, src_height=H)


def default_bands(H, n=48):
    """three bands of n rows: top edge, middle, bottom edge"""
    n = min(n, H)
    mid = max((H - n) // 2, 0)
    out = []
    for r0 in (0, mid, H - n):
        if (r0, r0 + n) not in out:
            out.append((r0, r0 + n))
    return out
