"""Pure-numpy models of the third-party arithmetic on the correct() path, and the
float32 "staged" chain the CUDA kernels are compared with bit for bit.

TEST INFRASTRUCTURE ONLY (see oracle/__init__.py).

The reference (radjkarl/imgProcessor 0.2.5) reaches this arithmetic through
  * ``scipy.ndimage.median_filter``          filters/medianThreshold.py:14,18
  * ``cv2.initUndistortRectifyMap``          camera/LensDistortion.py:355-357
  * ``cv2.remap(INTER_LINEAR, BORDER_CONSTANT)``  camera/LensDistortion.py:323-326
none of which is vendored under /root/reference.  The models below restate the
published algorithms (scipy ``mode='reflect'`` rank filter; OpenCV
``initUndistortRectifyMap`` Brown-Conrady evaluation in float64; OpenCV
``remap`` with INTER_BITS=5 fixed-point coordinates and the 32x32 bilinear
weight table) and are validated against scipy 1.18.1 / OpenCV 4.13.0 by
``tests/test_oracle_models.py``.
"""
import numpy as np

INTER_BITS = 5
INTER_TAB_SIZE = 1 << INTER_BITS          # 32
INTER_REMAP_COEF_BITS = 15                # u8 images use int16 weights scaled by 2**15
FLT_MAX = np.finfo(np.float32).max


# --------------------------------------------------------------------------
# median filter, scipy mode='reflect'  (d c b a | a b c d | d c b a)
# --------------------------------------------------------------------------
def reflect_index(i, n):
    """scipy.ndimage 'reflect' extension of index i into [0, n) (edge sample
    duplicated, period 2n).  Works for |i| >= n as well."""
    i = np.asarray(i)
    m = np.mod(i, 2 * n)
    return np.where(m >= n, 2 * n - 1 - m, m)


def median_filter_reflect(img, size=3):
    """size x size median, rank size*size//2, reflect border; dtype preserved.
    Model of ``scipy.ndimage.median_filter(img, size=size)`` as called at
    filters/medianThreshold.py:18."""
    img = np.asarray(img)
    assert img.ndim == 2
    h = size // 2
    H, W = img.shape
    rows = reflect_index(np.arange(-h, H + (size - 1 - h)), H)
    cols = reflect_index(np.arange(-h, W + (size - 1 - h)), W)
    pad = img[np.ix_(rows, cols)]
    stack = np.empty((size * size, H, W), dtype=img.dtype)
    k = 0
    for dy in range(size):
        for dx in range(size):
            stack[k] = pad[dy:dy + H, dx:dx + W]
            k += 1
    rank = (size * size) // 2
    stack.partition(rank, axis=0)
    return stack[rank].copy()


def median_threshold_model(img, threshold=0.1, size=3, condition='>'):
    """filters/medianThreshold.py:7-30 with the median modelled above.
    The median keeps img's dtype; the predicate is evaluated in float64
    (np.asfarray at :18).  Returns (new image (copy), mask) ; mask None if
    threshold <= 0."""
    img = np.asarray(img)
    if not threshold > 0:
        return img.copy(), None
    blur = median_filter_reflect(img, size).astype(np.float64)
    with np.errstate(divide='ignore', invalid='ignore', over='ignore'):
        rel = np.abs((img - blur) / blur)
        ind = rel > threshold if condition == '>' else rel < threshold
    out = img.copy()
    out[ind] = blur[ind]
    return out, ind


# --------------------------------------------------------------------------
# pointwise stage in float32 (what K1 must produce before the median)
# --------------------------------------------------------------------------
def pointwise_model(raw, dark=None, flat=None, nan_to_num=True, dark_ascent=None,
                    exposure_time=None, depth_bits=16):
    """float32( f64(raw) - f64(bg) [ / f64(flat) where flat != 0 ] ), then the
    float32 nan_to_num.  camera/CameraCalibration.py:408-410, 502, 507-516,
    525-526, 561.  The reference does the same in float64 and keeps float64;
    rounding the float64 value once to float32 is the 0-ulp target the
    north star's '<= 1 ulp' bar is measured from."""
    x = np.asarray(raw).astype(np.float64)
    if dark is not None:
        if dark_ascent is not None:
            # legacy tuple entry (:507-513): bg = offs + ascent*t, clipped to 2**depth-1
            bg = np.asarray(dark, np.float64) + np.asarray(dark_ascent, np.float64) * float(exposure_time)
            mx = float(2 ** depth_bits - 1)
            with np.errstate(invalid='ignore'):
                bg = np.where(bg > mx, mx, bg)
        else:
            bg = np.asarray(dark, dtype=np.float64)
        x = x - bg
    if flat is not None:
        f = np.asarray(flat, dtype=np.float64)
        nz = f != 0
        with np.errstate(over='ignore', invalid='ignore', divide='ignore'):
            x = np.where(nz, x / np.where(nz, f, 1.0), x)
    with np.errstate(over='ignore'):
        x32 = x.astype(np.float32)
    if nan_to_num:
        x32 = nan_to_num_f32(x32)
    return x32


def nan_to_num_f32(x):
    """np.nan_to_num in float32: NaN->0, +inf->FLT_MAX, -inf->-FLT_MAX
    (CameraCalibration.py:561 does it in float64, with DBL_MAX)."""
    return np.nan_to_num(np.asarray(x, np.float32))


# --------------------------------------------------------------------------
# Brown-Conrady map  (cv2.initUndistortRectifyMap, R = I, 5 coefficients)
# --------------------------------------------------------------------------
def undistort_map_model(K, dist, P, width, height):
    """float32 (mapx, mapy) of cv2.initUndistortRectifyMap(K, dist, None, P,
    (width,height), CV_32FC1): float64 evaluation of

        [X Y W]^T = P^-1 [u v 1]^T ; x = X/W ; y = Y/W ; r2 = x^2 + y^2
        kr = 1 + ((k3 r2 + k2) r2 + k1) r2
        xd = x kr + p1 2xy + p2 (r2 + 2x^2) ; yd = y kr + p1 (r2 + 2y^2) + p2 2xy
        mapx = fx xd + cx ; mapy = fy yd + cy

    then a cast to float32 (camera/LensDistortion.py:355-357)."""
    K = np.asarray(K, np.float64).reshape(3, 3)
    d = np.asarray(dist, np.float64).ravel()
    assert d.size == 5, 'only the 5-term model [k1,k2,p1,p2,k3] is on the path'
    k1, k2, p1, p2, k3 = d
    ir = np.linalg.inv(np.asarray(P, np.float64).reshape(3, 3)[:, :3])
    u = np.arange(width, dtype=np.float64)[None, :]
    v = np.arange(height, dtype=np.float64)[:, None]
    _x = u * ir[0, 0] + (v * ir[0, 1] + ir[0, 2])
    _y = u * ir[1, 0] + (v * ir[1, 1] + ir[1, 2])
    _w = u * ir[2, 0] + (v * ir[2, 1] + ir[2, 2])
    w = 1.0 / _w
    x = _x * w
    y = _y * w
    x2 = x * x
    y2 = y * y
    r2 = x2 + y2
    _2xy = 2 * x * y
    kr = 1 + ((k3 * r2 + k2) * r2 + k1) * r2
    xd = x * kr + p1 * _2xy + p2 * (r2 + 2 * x2)
    yd = y * kr + p1 * (r2 + 2 * y2) + p2 * _2xy
    mapx = K[0, 0] * xd + K[0, 2]
    mapy = K[1, 1] * yd + K[1, 2]
    return mapx.astype(np.float32), mapy.astype(np.float32)


# --------------------------------------------------------------------------
# cv2.remap, INTER_LINEAR, BORDER_CONSTANT
# --------------------------------------------------------------------------
def _cvround_x86(v):
    """cvRound / cvtps2dq on x86: round-half-even; NaN and out-of-int32-range
    give INT_MIN (0x80000000)."""
    v = np.asarray(v, np.float32)
    with np.errstate(invalid='ignore'):
        r = np.rint(v.astype(np.float64))
        bad = ~(np.abs(r) < 2147483648.0)            # also catches NaN
    r = np.where(bad, -2147483648.0, r)
    return r.astype(np.int64)


def fixed_point_coords(mapx, mapy):
    """OpenCV remap's conversion of float maps to 5-bit fixed point:
    sx = cvRound(mapx*32) (float32 product), ix = saturate_cast<short>(sx>>5),
    fx = sx & 31; same for y."""
    sx = _cvround_x86(np.asarray(mapx, np.float32) * np.float32(INTER_TAB_SIZE))
    sy = _cvround_x86(np.asarray(mapy, np.float32) * np.float32(INTER_TAB_SIZE))
    ix = np.clip(sx >> INTER_BITS, -32768, 32767)
    iy = np.clip(sy >> INTER_BITS, -32768, 32767)
    return ix, iy, sx & (INTER_TAB_SIZE - 1), sy & (INTER_TAB_SIZE - 1)


def bilinear_weights_f32(fx, fy):
    """The four float32 entries of OpenCV's BilinearTab_f[fy*32+fx]."""
    t = np.arange(INTER_TAB_SIZE, dtype=np.float32) * np.float32(1.0 / INTER_TAB_SIZE)
    one = np.float32(1)
    tx, ty = t[fx], t[fy]
    return (one - ty) * (one - tx), (one - ty) * tx, ty * (one - tx), ty * tx


def _gather(src, ix, iy, border, row0=0, height=None):
    """src holds rows [row0, row0 + src.shape[0]) of a frame of `height` rows (default: the whole frame)"""
    hb, W = src.shape
    H = hb if height is None else height
    ok = (ix >= 0) & (ix < W) & (iy >= 0) & (iy < H)
    local = iy - row0
    assert not (ok & ((local < 0) | (local >= hb))).any(), 'remap reads a source row outside the supplied band'
    v = src[np.clip(local, 0, hb - 1), np.clip(ix, 0, W - 1)]
    return np.where(ok, v, border)


def remap_model(src, mapx, mapy, border_value=0, src_row0=0, src_height=None):
    """Model of cv2.remap(src, mapx, mapy, INTER_LINEAR, BORDER_CONSTANT,
    border_value) for 2-D uint8 / uint16 / float32 / float64 images.
    * float32, uint16: float32 accumulate ((v00 w00 + v01 w01) + v10 w10) + v11 w11,
      every product and sum rounded separately (no FMA); uint16 then rounds
      half-even and saturates.
    * float64: same order in float64 with the float32 weights promoted.
    * uint8: int16 weights (w * 2**15), int32 accumulate, (acc + 2**14) >> 15.
    ``src`` may be a band of rows [src_row0, src_row0 + len(src)) of a frame of ``src_height`` rows (the maps stay
    in full-frame coordinates; every in-frame row the maps touch must lie inside the band)."""
    src = np.asarray(src)
    assert src.ndim == 2
    ix, iy, fx, fy = fixed_point_coords(mapx, mapy)
    w00, w01, w10, w11 = bilinear_weights_f32(fx, fy)
    H, W = src.shape
    if src_height is not None:
        H = src_height

    def gat(s, jx, jy, b):                          # band-aware gather
        return _gather(s, jx, jy, b, src_row0, H)
    # a 2x2 window entirely outside the image yields the border value itself (not a blend of it)
    outside = (ix >= W) | (ix + 1 < 0) | (iy >= H) | (iy + 1 < 0)
    if src.dtype == np.uint8:
        sc = float(1 << INTER_REMAP_COEF_BITS)
        iw = [np.rint(w.astype(np.float64) * sc).astype(np.int64) for w in (w00, w01, w10, w11)]
        s = src.astype(np.int64)
        b = int(np.clip(np.rint(border_value), 0, 255))
        acc = (gat(s, ix, iy, b) * iw[0] + gat(s, ix + 1, iy, b) * iw[1]
               + gat(s, ix, iy + 1, b) * iw[2] + gat(s, ix + 1, iy + 1, b) * iw[3])
        r = (acc + (1 << (INTER_REMAP_COEF_BITS - 1))) >> INTER_REMAP_COEF_BITS
        return np.where(outside, b, np.clip(r, 0, 255)).astype(np.uint8)
    if src.dtype == np.float64:
        acc_t = np.float64
        b = np.float64(border_value)
    elif src.dtype == np.uint16:
        acc_t = np.float32
        b = np.float32(np.clip(np.rint(border_value), 0, 65535))
    else:
        acc_t = np.float32
        b = np.float32(border_value)
    s = src.astype(acc_t)
    w00, w01, w10, w11 = (w.astype(acc_t) for w in (w00, w01, w10, w11))
    with np.errstate(over='ignore', invalid='ignore'):
        r = ((gat(s, ix, iy, b) * w00 + gat(s, ix + 1, iy, b) * w01)
             + gat(s, ix, iy + 1, b) * w10) + gat(s, ix + 1, iy + 1, b) * w11
    r = np.where(outside, b, r)
    if src.dtype == np.uint16:
        return np.clip(np.rint(r), 0, 65535).astype(np.uint16)
    return r.astype(src.dtype)


# --------------------------------------------------------------------------
# the float32 staged chain = bit-exact target of K1 -> K2
# --------------------------------------------------------------------------
def correct_chain_f32(raw, dark=None, flat=None, threshold=0.1, size=3, lens=None,
                      border_value=0.0, mapxy=None):
    """dark/flat in float64 registers rounded once to float32, float32 median +
    float64 predicate, analytic float64 map cast to float32, OpenCV fixed-point
    bilinear in float32.  ``lens`` = (K, dist, P) or None; ``mapxy`` overrides
    the analytic map (e.g. with cv2's own maps)."""
    do_med = threshold is not None and threshold > 0
    x = pointwise_model(raw, dark, flat, nan_to_num=do_med)
    mask = None
    if do_med:
        x, mask = median_threshold_model(x, threshold, size)
    if lens is not None or mapxy is not None:
        H, W = x.shape
        if mapxy is None:
            K, dist, P = lens
            mapxy = undistort_map_model(K, dist, P, W, H)
        x = remap_model(x, mapxy[0], mapxy[1], border_value)
    return x, mask
