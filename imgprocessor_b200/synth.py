"""Seeded synthetic inputs for the correct() path: an EL-like
scene with hot/dead pixels, a dark-current map, a Kang-Weiss flat field and
5-coefficient lenses.  numpy only; ``scene_torch`` builds the same kind of frame
directly on a CUDA device for bench.py (content cannot affect timing much: the
kernels are select-based, only the rare float64 predicate branch is data dependent).

The flat field uses the off-axis term of the Kang-Weiss model the reference
ships as imgProcessor/equations/vignetting.py:9-37 (A = 1/(1+(r/f)^2)^2).
"""
import numpy as np

# lens presets: (fx, fy, cx, cy, k1, k2, k3, p1, p2) as LensDistortion.setCameraParams
# takes them (camera/LensDistortion.py:373-380)


def lens_moderate(H, W):
    return (float(W), float(W), W / 2.0, H / 2.0, -0.2, 0.05, 0.0, 1e-3, -1e-3)


def lens_strong(H, W):
    s = W / 8192.0
    return (8192.0 * s, 8192.0 * s, 4100.0 * s, 4090.0 * s * H / W, -0.25, 0.08, -0.01, 2e-3, -1.5e-3)


# coefficients cv2.calibrateCamera gives on the 30 fixture photos shipped under
# imgProcessor/media/lens_distortion, image shape (501, 665)
LENS_REALISTIC_SHAPE = (501, 665)


def lens_realistic():
    return (1550.577, 1547.862, 317.596, 214.359, -0.08870538, 0.25284511, -2.78044631,
            0.00350456, 0.00326603)


def camera_matrix(params):
    fx, fy, cx, cy = params[:4]
    K = np.zeros((3, 3))
    K[0, 0], K[1, 1], K[0, 2], K[1, 2], K[2, 2] = fx, fy, cx, cy, 1.0
    return K


def dist_coeffs(params):
    k1, k2, k3, p1, p2 = params[4:]
    return np.array([[k1, k2, p1, p2, k3]], dtype=np.float64)


def _smooth_field(H, W, rng, cell=32):
    """low-resolution uniform noise, bilinearly upsampled (a cheap stand-in for
    a sigma~8px Gaussian blur of white noise), rescaled to [0, 1]."""
    gh, gw = H // cell + 3, W // cell + 3
    g = rng.random((gh, gw))
    ys = (np.arange(H) + 0.5) / cell
    xs = (np.arange(W) + 0.5) / cell
    y0 = ys.astype(np.int64)
    x0 = xs.astype(np.int64)
    fy = (ys - y0)[:, None]
    fx = (xs - x0)[None, :]
    a = g[np.ix_(y0, x0)]
    b = g[np.ix_(y0, x0 + 1)]
    c = g[np.ix_(y0 + 1, x0)]
    d = g[np.ix_(y0 + 1, x0 + 1)]
    f = (a * (1 - fx) + b * fx) * (1 - fy) + (c * (1 - fx) + d * fx) * fy
    f -= f.min()
    f /= max(f.max(), 1e-12)
    return f


def defect_masks(H, W, seed, p_hot=1e-3, p_dead=5e-4):
    rng = np.random.default_rng(seed + 7919)
    r = rng.random((H, W))
    return r < p_hot, (r >= p_hot) & (r < p_hot + p_dead)


def scene(H, W, seed, dtype=np.uint16, full_scale=None, hot_dead=True):
    """EL-like frame: smooth field in [0.2,0.9]*FS, dark cell-grid lines every 256 px,
    0.5 % FS Gaussian read noise, hot pixels -> FS, dead pixels -> 0."""
    dtype = np.dtype(dtype)
    if full_scale is None:
        full_scale = {np.dtype(np.uint16): 65535.0, np.dtype(np.uint8): 255.0}.get(dtype, 4095.0)
    rng = np.random.default_rng(seed)
    img = (0.2 + 0.7 * _smooth_field(H, W, rng)) * full_scale
    img[::256, :] *= 0.35
    img[:, ::256] *= 0.35
    img += rng.normal(0.0, 0.005 * full_scale, (H, W))
    if hot_dead:
        hot, dead = defect_masks(H, W, seed)
        img[hot] = full_scale
        img[dead] = 0.0
    img = np.clip(img, 0.0, full_scale)
    if dtype.kind in 'ui':
        img = np.rint(img)
    return img.astype(dtype)


def dark_map(H, W, seed=11):
    rng = np.random.default_rng(seed)
    d = 100.0 + 10.0 * rng.random((H, W))
    hot, _ = defect_masks(H, W, seed)
    d[hot] += 500.0
    return d.astype(np.float32)


def flat_map(H, W, seed=13, p_zero=1e-5):
    rng = np.random.default_rng(seed)
    y, x = np.mgrid[0:H, 0:W].astype(np.float64)
    r2 = (x - (W / 2.0 + 7.0)) ** 2 + (y - (H / 2.0 - 5.0)) ** 2
    f = 1.0 / (1.0 + r2 / (0.9 * W) ** 2) ** 2
    f *= rng.normal(1.0, 0.01, (H, W))
    f = np.clip(f, 0.05, None)
    f[rng.random((H, W)) < p_zero] = 0.0
    return f.astype(np.float32)


def scene_torch(n, H, W, seed, device, dtype='uint16'):
    """(n,H,W) frames of the same kind generated on ``device`` with torch ops
    (bench.py; avoids minutes of host RNG for 256 frames)."""
    import torch
    import torch.nn.functional as F
    g = torch.Generator(device=device)
    g.manual_seed(int(seed))
    fs = 65535.0 if dtype == 'uint16' else 4095.0
    out_dtype = torch.uint16 if dtype == 'uint16' else torch.float32
    out = torch.empty((n, H, W), dtype=out_dtype, device=device)
    for i in range(n):
        low = torch.rand((1, 1, H // 32 + 3, W // 32 + 3), generator=g, device=device)
        f = F.interpolate(low, size=(H, W), mode='bilinear', align_corners=False)[0, 0]
        img = (0.2 + 0.7 * f) * fs
        img[::256, :] *= 0.35
        img[:, ::256] *= 0.35
        img += torch.randn((H, W), generator=g, device=device) * (0.005 * fs)
        r = torch.rand((H, W), generator=g, device=device)
        img = torch.where(r < 1e-3, torch.full_like(img, fs), img)
        img = torch.where((r >= 1e-3) & (r < 1.5e-3), torch.zeros_like(img), img)
        img = img.clamp_(0.0, fs)
        if dtype == 'uint16':
            out[i] = img.round_().to(torch.int32).to(torch.uint16)
        else:
            out[i] = img
    return out
