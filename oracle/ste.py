"""Single-time-effect (STE) removal: the multi-image branch of CameraCalibration.correct()
(SURVEY §8 rows a11 / f1).

TEST INFRASTRUCTURE ONLY (see oracle/__init__.py).

Reference (paths relative to /root/reference/imgProcessor/):
  * call site                          camera/CameraCalibration.py:385-406, 484-498
  * SingleTimeEffectDetection          features/SingleTimeEffectDetection.py:23-75
  * removeSinglePixels (numba)         filters/removeSinglePixels.py:4-33
  * NoiseLevelFunction.boundedFunction camera/NoiseLevelFunction.py:94-107

PARITY UNPINNED for one ingredient: the running average lives in
``fancytools.math.MaskedMovingAverage`` (fancytools is an un-vendored, un-pinned dependency,
setup.py:34-41, and is absent from this environment).  Its published algorithm — a per-pixel count
``n`` and the incremental mean ``avg += (x - avg) / n`` applied where the mask is set (everywhere
without a mask) — is restated in ``MaskedMovingAverage`` below.  Everything else is pinned:
tests/golden/ste.npz holds outputs of the reference's own SingleTimeEffectDetection, removeSinglePixels
and boundedFunction, executed unmodified with this restated class injected for the missing module.
"""
import numpy as np


class MaskedMovingAverage(object):
    """restatement of fancytools.math.MaskedMovingAverage (calcVariance=False): see module docstring"""

    def __init__(self, shape, calcVariance=False, dtype=float):
        if calcVariance:
            raise NotImplementedError('variance tracking is not on the correct() path')
        self.n = np.zeros(shape, dtype=int)
        self.avg = np.zeros(shape, dtype=dtype)

    def update(self, data, mask=None):
        if mask is None:
            self.n += 1
            self.avg += (data - self.avg) / self.n
        else:
            self.n[mask] += 1
            a = self.avg[mask]
            self.avg[mask] = a + (data[mask] - a) / self.n[mask]


def bounded_function(x, minY, ax, ay):
    """NoiseLevelFunction.boundedFunction (:94-107): max(nan_to_num(ay * sqrt(x - ax)), minY)"""
    with np.errstate(invalid='ignore'):
        y = ay * (x - ax) ** 0.5
    return np.maximum(np.nan_to_num(y), minY)


def remove_single_pixels(mask):
    """removeSinglePixels (:4-33), vectorised: a True pixel without a True 8-neighbour becomes False.  (The in-place
    raster scan of the reference gives the same result: a pixel is only removed when it has no neighbour at all, so
    no removal can change another pixel's verdict.)"""
    m = np.asarray(mask, bool)
    p = np.pad(m, 1).astype(np.int32)
    H, W = m.shape
    cnt = sum(p[1 + dy:1 + dy + H, 1 + dx:1 + dx + W] for dy in (-1, 0, 1) for dx in (-1, 0, 1) if (dy, dx) != (0, 0))
    return m & (cnt > 0)


def ste_average(images, nlf_params, nStd=4, return_mask=False):
    """SingleTimeEffectDetection(images, nStd=nStd, noise_level_function=boundedFunction(., *nlf_params)).noSTE
    (float64), optionally with the accumulated STE mask (save_ste_indices=True)."""
    i1 = np.asarray(images[0], np.float64)
    i2 = np.asarray(images[1])
    mma = MaskedMovingAverage(i1.shape)
    mma.update(np.min((i1, i2), axis=0))
    thr = bounded_function(mma.avg, *nlf_params) * nStd
    ste_any = np.zeros(i1.shape, bool)
    for img in [np.max((i1, i2), axis=0)] + [np.asarray(i) for i in images[2:]]:
        ste = remove_single_pixels((img - mma.avg) > thr)
        mma.update(img, ~ste)
        ste_any |= ste
    return (mma.avg, ste_any) if return_mask else mma.avg
