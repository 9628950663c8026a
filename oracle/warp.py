"""Pure-numpy model of ``cv2.warpPerspective`` as PerspectiveCorrection uses it
(SURVEY §8 row f3).

TEST INFRASTRUCTURE ONLY (see oracle/__init__.py).

Reference call sites (paths relative to /root/reference/imgProcessor/):
  * ``PerspectiveCorrection.correct``    camera/PerspectiveCorrection.py:380-406
        optional ``img / tiltFactor`` then
        ``cv2.warpPerspective(img, h, newBorders[::-1], flags=cv2.INTER_LANCZOS4)``
  * ``PerspectiveCorrection.uncorrect``  camera/PerspectiveCorrection.py:374-378
        ``cv2.warpPerspective(img, h, s[::-1], flags=INTER_CUBIC | WARP_INVERSE_MAP)``
  * ``homography`` from a quad           camera/PerspectiveCorrection.py:147-149
        ``cv2.getPerspectiveTransform`` (host-side setup, stays on OpenCV)

OpenCV is un-vendored; the model restates its published algorithm (imgwarp.cpp):
  * ``cv::invert`` of the 3x3 matrix by cofactors unless WARP_INVERSE_MAP,
  * per output pixel, in float64, blockwise (``bw0`` columns per block):
        X0 = M0*bx + M1*y + M2 ; W = W0 + M6*x1 ; W = W ? 32/W : 0
        X  = cvRound(clamp(INT)((X0 + M0*x1) * W))      -> ix = X >> 5, fx = X & 31
  * Lanczos4 (8x8 taps) / bicubic (4x4 taps, A = -0.75) float32 coefficient tables
    for the 32 sub-pixel phases, 2-D weight = float32(wy * wx),
  * interior pixels: one left-to-right sum per tap row, the row sums added top to bottom
    (OpenCV 4.13.0 does this for bicubic as well), float32 accumulation (float64 for float64 images), no FMA;
    border pixels: ``cv + sum((S - cv) * w)`` over the in-range taps in raster order;
    window entirely outside: the border value.
Validated bit for bit against OpenCV 4.13.0 in tests/test_oracle_models.py and
against the reference's own PerspectiveCorrection outputs (tests/golden).
uint8 images use OpenCV's fixed-point branch: int16 weights = saturate_cast<short>(wy * wx * 2**15) with the sum of each
phase pair's table forced to 2**15 on one of the four central taps (initInterTab2D), int32 accumulation, rounding shift.
"""
import numpy as np

INTER_TAB_SIZE = 32


def invert3x3_cv(M):
    """cv::invert(DECOMP_LU) for a 3x3 float64 matrix: cofactors times 1/det."""
    m = np.asarray(M, np.float64).reshape(3, 3)
    d = (m[0, 0] * (m[1, 1] * m[2, 2] - m[1, 2] * m[2, 1])
         - m[0, 1] * (m[1, 0] * m[2, 2] - m[1, 2] * m[2, 0])
         + m[0, 2] * (m[1, 0] * m[2, 1] - m[1, 1] * m[2, 0]))
    if d == 0:
        return np.zeros((3, 3))
    d = 1.0 / d
    t = np.empty(9)
    t[0] = (m[1, 1] * m[2, 2] - m[1, 2] * m[2, 1]) * d
    t[1] = (m[0, 2] * m[2, 1] - m[0, 1] * m[2, 2]) * d
    t[2] = (m[0, 1] * m[1, 2] - m[0, 2] * m[1, 1]) * d
    t[3] = (m[1, 2] * m[2, 0] - m[1, 0] * m[2, 2]) * d
    t[4] = (m[0, 0] * m[2, 2] - m[0, 2] * m[2, 0]) * d
    t[5] = (m[0, 2] * m[1, 0] - m[0, 0] * m[1, 2]) * d
    t[6] = (m[1, 0] * m[2, 1] - m[1, 1] * m[2, 0]) * d
    t[7] = (m[0, 1] * m[2, 0] - m[0, 0] * m[2, 1]) * d
    t[8] = (m[0, 0] * m[1, 1] - m[0, 1] * m[1, 0]) * d
    return t.reshape(3, 3)


def lanczos4_table():
    """cv::interpolateLanczos4 for the 32 phases: float32 [32][8]."""
    s45 = 0.70710678118654752440084436210485
    cs = ((1, 0), (-s45, -s45), (0, 1), (s45, -s45), (-1, 0), (s45, s45), (0, -1), (-s45, s45))
    tab = np.zeros((INTER_TAB_SIZE, 8), np.float32)
    tab[0, 3] = 1                                   # x < FLT_EPSILON: unit tap
    f32 = np.float32
    for k in range(1, INTER_TAB_SIZE):
        x = f32(k) * f32(1.0 / INTER_TAB_SIZE)
        y0 = -(np.float64(x) + 3) * np.pi * 0.25
        s0, c0 = np.sin(y0), np.cos(y0)
        co = np.zeros(8, np.float32)
        total = f32(0)
        for i in range(8):
            yf = f32(f32(x + f32(3)) - f32(i))      # -(x+3-i) is float arithmetic in OpenCV
            y = -np.float64(yf) * np.pi * 0.25
            co[i] = f32((cs[i][0] * s0 + cs[i][1] * c0) / (y * y))
            total = f32(total + co[i])
        total = f32(f32(1) / total)
        tab[k] = co * total
    return tab


def cubic_table():
    """cv::interpolateCubic (A = -0.75) for the 32 phases: float32 [32][4]."""
    f32 = np.float32
    A = f32(-0.75)
    tab = np.zeros((INTER_TAB_SIZE, 4), np.float32)
    for k in range(INTER_TAB_SIZE):
        x = f32(k) * f32(1.0 / INTER_TAB_SIZE)
        x1 = f32(x + f32(1))
        c0 = f32(f32(f32(f32(f32(f32(A * x1) - f32(f32(5) * A)) * x1) + f32(f32(8) * A)) * x1) - f32(f32(4) * A))
        c1 = f32(f32(f32(f32(f32(f32(A + f32(2)) * x) - f32(A + f32(3))) * x) * x) + f32(1))
        xm = f32(f32(1) - x)
        c2 = f32(f32(f32(f32(f32(f32(A + f32(2)) * xm) - f32(A + f32(3))) * xm) * xm) + f32(1))
        c3 = f32(f32(f32(f32(1) - c0) - c1) - c2)
        tab[k] = (c0, c1, c2, c3)
    return tab


def fixed_point_table(tab):
    """initInterTab2D, fixed-point branch: int16 weights [32][32][n*n] (phase fy, phase fx, tap row-major)."""
    n = tab.shape[1]
    out = np.zeros((INTER_TAB_SIZE, INTER_TAB_SIZE, n * n), np.int16)
    k2 = n // 2
    for i in range(INTER_TAB_SIZE):
        for j in range(INTER_TAB_SIZE):
            v = (tab[i][:, None] * tab[j][None, :]).astype(np.float32)
            it = np.clip(np.rint(v * np.float32(32768)), -32768, 32767).astype(np.int64).ravel()
            diff = int(it.sum()) - 32768
            if diff != 0:
                Mk = mk = (k2, k2)
                for a in range(k2, k2 + 2):
                    for b in range(k2, k2 + 2):
                        if it[a * n + b] < it[mk[0] * n + mk[1]]:
                            mk = (a, b)
                        elif it[a * n + b] > it[Mk[0] * n + Mk[1]]:
                            Mk = (a, b)
                k = Mk if diff < 0 else mk
                it[k[0] * n + k[1]] -= diff
            out[i, j] = it
    return out


def _warp_u8(src, tab, n, off, coords, border_value):
    it = fixed_point_table(tab).astype(np.int64)
    ix, iy, fx, fy = coords
    H, W = src.shape
    sx, sy = ix - off, iy - off
    fast = (sx >= 0) & (sx < max(W - (n - 1), 0)) & (sy >= 0) & (sy < max(H - (n - 1), 0))
    outside = (sx >= W) | (sx + n <= 0) | (sy >= H) | (sy + n <= 0)
    w = it[fy, fx]
    cv = int(np.clip(np.rint(border_value), 0, 255))
    total = np.zeros(ix.shape, np.int64)
    edge = np.full(ix.shape, cv * 32768, np.int64)
    s64 = src.astype(np.int64)
    for r in range(n):
        yy = sy + r
        yin = (yy >= 0) & (yy < H)
        yc = np.clip(yy, 0, H - 1)
        for c in range(n):
            xx = sx + c
            xin = (xx >= 0) & (xx < W)
            v = s64[yc, np.clip(xx, 0, W - 1)]
            total += v * w[..., r * n + c]
            edge += np.where(yin & xin, (v - cv) * w[..., r * n + c], 0)
    res = np.where(fast, total, np.where(outside, cv * 32768, edge))
    return np.clip((res + (1 << 14)) >> 15, 0, 255).astype(np.uint8)


def warp_coords(M, dsize, inverse_map=False):
    """Fixed-point source coordinates of every output pixel: (ix, iy, fx, fy), int64 [h][w]."""
    w, h = int(dsize[0]), int(dsize[1])
    M = np.asarray(M, np.float64).reshape(3, 3)
    if not inverse_map:
        M = invert3x3_cv(M)
    M = M.ravel()
    bh0 = min(16, h)
    bw0 = min(1024 // bh0, w)
    xs = np.arange(w)
    bx = ((xs // bw0) * bw0)[None].astype(np.float64)
    x1 = (xs[None] - bx).astype(np.float64)
    ys = np.arange(h)[:, None].astype(np.float64)
    X0 = M[0] * bx + M[1] * ys + M[2]
    Y0 = M[3] * bx + M[4] * ys + M[5]
    W0 = M[6] * bx + M[7] * ys + M[8]
    W = W0 + M[6] * x1
    with np.errstate(divide='ignore', invalid='ignore', over='ignore'):
        W = np.where(W != 0, INTER_TAB_SIZE / W, 0.0)
        fX = (X0 + M[0] * x1) * W
        fY = (Y0 + M[3] * x1) * W
    imin, imax = -2147483648.0, 2147483647.0
    # std::max(INT_MIN, std::min(INT_MAX, v)) : a NaN ends up as INT_MAX
    fX = np.where(np.isnan(fX), imax, np.clip(fX, imin, imax))
    fY = np.where(np.isnan(fY), imax, np.clip(fY, imin, imax))
    X = np.rint(fX).astype(np.int64)
    Y = np.rint(fY).astype(np.int64)
    ix = np.clip(X >> 5, -32768, 32767)
    iy = np.clip(Y >> 5, -32768, 32767)
    return ix, iy, X & 31, Y & 31


def warp_perspective_model(src, M, dsize, interpolation='lanczos4', inverse_map=False, border_value=0.0,
                           divide_by=None):
    """cv2.warpPerspective(src, M, dsize, flags=INTER_LANCZOS4 | INTER_CUBIC [| WARP_INVERSE_MAP],
    borderMode=BORDER_CONSTANT, borderValue) for a 2-D uint8 / uint16 / float32 / float64 image.
    divide_by: the tilt-factor division of PerspectiveCorrection.correct (:394-400) applied first,
    in float64 (the image becomes float64, as np.asfarray does there)."""
    src = np.asarray(src)
    if divide_by is not None:
        src = np.asarray(src, np.float64) / np.asarray(divide_by, np.float64)
    if src.dtype not in (np.uint8, np.uint16, np.float32, np.float64):
        raise TypeError('unmodelled dtype %s' % src.dtype)
    if interpolation == 'lanczos4':
        tab, n, off = lanczos4_table(), 8, 3
    elif interpolation == 'cubic':
        tab, n, off = cubic_table(), 4, 1
    else:
        raise ValueError(interpolation)
    ix, iy, fx, fy = warp_coords(M, dsize, inverse_map)
    if src.dtype == np.uint8:
        return _warp_u8(src, tab, n, off, (ix, iy, fx, fy), border_value)
    H, W = src.shape
    wt = np.float64 if src.dtype == np.float64 else np.float32
    sx, sy = ix - off, iy - off
    fast = (sx >= 0) & (sx < max(W - (n - 1), 0)) & (sy >= 0) & (sy < max(H - (n - 1), 0))
    outside = (sx >= W) | (sx + n <= 0) | (sy >= H) | (sy + n <= 0)
    wy, wx = tab[fy], tab[fx]
    cv = wt(border_value)
    if src.dtype == np.uint16:
        cv = wt(np.uint16(np.clip(np.rint(border_value), 0, 65535)))
    total = np.zeros(ix.shape, wt)                  # interior arithmetic
    edge = np.full(ix.shape, cv * wt(1), wt)        # border arithmetic
    acc = None
    with np.errstate(over='ignore', invalid='ignore'):
        for r in range(n):
            yy = sy + r
            yin = (yy >= 0) & (yy < H)
            yc = np.clip(yy, 0, H - 1)
            acc = None                              # one left-to-right sum per tap row
            for c in range(n):
                xx = sx + c
                xin = (xx >= 0) & (xx < W)
                xc = np.clip(xx, 0, W - 1)
                w2 = (wy[..., r] * wx[..., c]).astype(np.float32).astype(wt)
                v = src[yc, xc].astype(wt)
                p = (v * w2).astype(wt)
                acc = p if acc is None else (acc + p).astype(wt)
                e = ((v - cv).astype(wt) * w2).astype(wt)
                edge = np.where(yin & xin, (edge + e).astype(wt), edge)
            total = acc if r == 0 else (total + acc).astype(wt)
    res = np.where(fast, total, np.where(outside, cv, edge))
    if src.dtype == np.uint16:
        with np.errstate(invalid='ignore'):
            r = np.rint(res.astype(np.float64))
        return np.where(np.isnan(r), 0, np.clip(r, 0, 65535)).astype(np.uint16)
    return res.astype(src.dtype)
