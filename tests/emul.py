"""ctypes access to the host emulation of the kernels' per-pixel arithmetic
(tests/host_emul/emul.cpp, compiled from imgprocessor_b200/csrc/imgcorr_core.cuh with g++)."""
import ctypes
import os
import subprocess

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
SRC = os.path.join(HERE, 'host_emul', 'emul.cpp')
CORE = os.path.join(os.path.dirname(HERE), 'imgprocessor_b200', 'csrc', 'imgcorr_core.cuh')
STE = os.path.join(os.path.dirname(HERE), 'imgprocessor_b200', 'csrc', 'imgcorr_ste.cuh')
WARP = os.path.join(os.path.dirname(HERE), 'imgprocessor_b200', 'csrc', 'imgcorr_warp.cuh')
NET = os.path.join(os.path.dirname(HERE), 'imgprocessor_b200', 'csrc', 'median25_net.inc')
OUT = os.path.join(HERE, '_build', 'libimgcorr_emul.so')
_DT = {np.dtype(np.uint8): 0, np.dtype(np.uint16): 1, np.dtype(np.float32): 2, np.dtype(np.float64): 3}
_lib = None


def lib():
    global _lib
    if _lib is None:
        os.makedirs(os.path.dirname(OUT), exist_ok=True)
        if not os.path.exists(OUT) or os.path.getmtime(OUT) < max(os.path.getmtime(SRC), os.path.getmtime(CORE), os.path.getmtime(NET), os.path.getmtime(WARP), os.path.getmtime(STE)):
            subprocess.check_call(['g++', '-O2', '-std=c++17', '-ffp-contract=off', '-fPIC', '-shared',
                                   '-fvisibility=hidden', '-o', OUT, SRC, '-lm'])
        _lib = ctypes.CDLL(OUT)
    return _lib


def _p(a):
    return None if a is None else a.ctypes.data_as(ctypes.c_void_p)


def k1(raw, dark=None, flat=None, threshold=0.1, ksize=3, cond='>', flags=7, out_dtype=None, ascent=None,
       exposure=0.0, depth_bits=16, want_mask=True):
    raw = np.ascontiguousarray(raw)
    H, W = raw.shape
    f32 = lambda a: None if a is None else np.ascontiguousarray(a, np.float32)
    dark, flat, ascent = f32(dark), f32(flat), f32(ascent)
    if out_dtype is None:
        out_dtype = np.float64 if raw.dtype == np.float64 else np.float32
    out = np.empty((H, W), out_dtype)
    mask = np.zeros((H, W), np.uint8) if want_mask else None
    lib().emul_k1(_p(raw), _DT[raw.dtype], _p(dark), _p(ascent), _p(flat), _p(out), _DT[np.dtype(out_dtype)], _p(mask),
                  H, W, ctypes.c_double(threshold), ksize, 0 if cond == '>' else 1, flags,
                  ctypes.c_double(exposure), ctypes.c_double(2.0 ** depth_bits - 1))
    return out, (mask.astype(bool) if want_mask else None)


def maps(K, dist, P, H, W):
    K = np.ascontiguousarray(K, np.float64)
    d = np.ascontiguousarray(np.asarray(dist, np.float64).ravel())
    P = np.ascontiguousarray(np.asarray(P, np.float64)[:, :3])
    mx = np.empty((H, W), np.float32)
    my = np.empty((H, W), np.float32)
    r = lib().emul_maps(_p(K), _p(d), _p(P), H, W, _p(mx), _p(my))
    assert r == 0
    return mx, my


def remap(src, mapx, mapy, border=0.0, window=None, widen=False):
    src = np.ascontiguousarray(src)
    H, W = src.shape
    x0, y0, ow, oh = window if window is not None else (0, 0, W, H)
    dt = _DT[src.dtype]
    out_dtype = src.dtype
    if widen:
        assert src.dtype == np.float32
        dt, out_dtype = 4, np.float64
    dst = np.empty((oh, ow), out_dtype)
    lib().emul_remap(_p(src), dt, _p(dst), H, W, _p(np.ascontiguousarray(mapx, np.float32)),
                     _p(np.ascontiguousarray(mapy, np.float32)), ctypes.c_double(border), x0, y0, ow, oh)
    return dst


def warp_tables(M=None):
    """(lanczos4 [32][8], cubic [32][4], inverse of M by the cv::invert formula)"""
    lz = np.empty((32, 8), np.float32)
    cu = np.empty((32, 4), np.float32)
    M = np.ascontiguousarray(np.eye(3) if M is None else M, np.float64)
    Mi = np.empty((3, 3), np.float64)
    lib().emul_warp_tables(_p(lz), _p(cu), _p(M), _p(Mi))
    return lz, cu, Mi


def warp(src, M, dsize, interpolation='lanczos4', inverse_map=False, border=0.0):
    src = np.ascontiguousarray(src)
    H, W = src.shape
    dw, dh = int(dsize[0]), int(dsize[1])
    dst = np.empty((dh, dw), src.dtype)
    r = lib().emul_warp(_p(src), _DT[src.dtype], _p(dst), H, W, dh, dw, _p(np.ascontiguousarray(M, np.float64)),
                        {'lanczos4': 4, 'cubic': 2}[interpolation], int(bool(inverse_map)), ctypes.c_double(border))
    assert r == 0
    return dst


def ste(frames, nlf, n_std=4.0, want_mask=False):
    frames = np.ascontiguousarray(frames, np.float64)
    n, H, W = frames.shape
    avg = np.empty((H, W), np.float64)
    mask = np.zeros((H, W), np.uint8) if want_mask else None
    lib().emul_ste(_p(frames), n, H, W, _p(np.ascontiguousarray(nlf, np.float64)), ctypes.c_double(n_std), _p(avg), _p(mask))
    return (avg, mask.astype(bool)) if want_mask else avg


def keys(x):
    x = np.ascontiguousarray(x, np.float32)
    k = np.empty(x.shape, np.int32)
    back = np.empty(x.shape, np.float32)
    lib().emul_keys(_p(x), x.size, _p(k), _p(back))
    return k, back
