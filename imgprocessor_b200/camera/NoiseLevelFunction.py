"""Noise level function sigma(intensity) — the part of imgProcessor.camera.NoiseLevelFunction the multi-exposure
branch of CameraCalibration.correct() reaches.

Reference (paths relative to /root/reference/imgProcessor/):
  * boundedFunction / function           camera/NoiseLevelFunction.py:94-107
  * oneImageNLF                          camera/NoiseLevelFunction.py:153-159   called from
                                         features/SingleTimeEffectDetection.py:43-45 when correct() has no 'noise'
                                         calibration (camera/CameraCalibration.py:392-406)
  * calcNLF (binned average absolute deviation of image - median) :176-271, _getMinMax :162-173
  * _evaluate / _validI / _fit / smooth (square-root fit, polynomial fallback) :73-150

This is host logic that runs ONCE per calibration object (the fitted function is kept in
``CameraCalibration.noise_level_function``, as in the reference); the only image-sized operation besides the binning
is the 3x3 median, which runs on the GPU through kernel K1 (``Engine.median3x3``).  The fit itself is
``scipy.optimize.curve_fit`` exactly as in the reference.  Same names, arguments and return values as the reference
functions it mirrors; written from their behaviour, not copied.
"""
import numpy as np

RMS_PER_AAD = (2.0 / np.pi) ** -0.5          # average absolute deviation -> RMS for Gaussian noise
MEDIAN_NOISE_GAIN = 1.0 + 1.0 / 3 ** 2       # img - median3x3(img) carries 1 + 1/9 of the noise variance scale used by the reference
DEFAULT_BINS = 100


def function(x, ax, ay):
    """ay * sqrt(x - ax); NaN left of ax (as numpy's ``** 0.5``)"""
    with np.errstate(invalid='ignore'):
        return ay * (x - ax) ** 0.5


def boundedFunction(x, minY, ax, ay):
    """function() bounded from below by minY; NaN (x < ax) counts as 0 before the bound"""
    return np.maximum(np.nan_to_num(function(x, ax, ay)), minY)


class FittedNLF(object):
    """the callable oneImageNLF returns.  ``params`` = (minY, ax, ay) when the square-root model could be fitted (then the
    GPU evaluates it per pixel inside K4), else None and ``poly`` / ``xrange`` / ``const`` describe the fallback."""

    def __init__(self, params=None, poly=None, xrange=None, const=None):
        self.params = None if params is None else tuple(float(v) for v in params)
        self.poly, self.xrange, self.const = poly, xrange, const

    def __call__(self, x):
        if self.params is not None:
            return boundedFunction(x, *self.params)
        if self.poly is not None:
            return np.poly1d(self.poly)(np.clip(x, self.xrange[0], self.xrange[1]))
        return self.const


def _median3x3(img):
    from .. import engine as _engine
    h, w = img.shape
    return _engine.get_engine(h, w).median3x3(img)


def _getMinMax(img):
    """intensity range holding most pixels: mean +- 3 sigma, clipped to the data range and to >= 0"""
    av, sd = np.mean(img), np.std(img)
    return max(img.min(), av - 3 * sd, 0), min(img.max(), av + 3 * sd)


def calcNLF(img, img2=None, signal=None, mn_mx_nbins=None, x=None, averageFn='AAD', signalFromMultipleImages=False):
    """binned noise estimate: returns (x, y, weights, signal) — bin centres, noise (RMS-scaled) per bin, samples per bin
    and the signal image used for binning.  One image: noise = (img - median3x3(img)) * (1 + 1/9); two images of the same
    scene: noise = (img - img2) / sqrt(2), signal = median3x3 of their mean."""
    if averageFn == 'AAD':
        def average(d):
            return np.mean(np.abs(d)) * RMS_PER_AAD
    else:
        def average(d):
            return (d ** 2).mean() ** 0.5
    img = np.asarray(img, dtype=np.float64)
    if img2 is None:
        if signal is None:
            signal = _median3x3(img)
        noise = img - signal
        if not signalFromMultipleImages:
            noise = noise * MEDIAN_NOISE_GAIN
    else:
        img2 = np.asarray(img2, dtype=np.float64)
        noise = img - img2
        noise /= 2 ** 0.5
        if signal is None:
            signal = _median3x3(0.5 * (img + img2))
    if mn_mx_nbins is not None:
        lo, hi, nbins = mn_mx_nbins
        min_len = 0
    else:
        lo, hi = _getMinMax(signal)
        min_len = int(img.shape[0] * img.shape[1] * 1e-3)
        if min_len < 1:
            min_len = 5
        nbins = DEFAULT_BINS
        if hi - lo < nbins:
            nbins = int(hi - lo)
    step = (hi - lo) / nbins
    y = np.full(nbins, np.nan)
    weights = np.zeros(nbins)
    fill_x = x is None
    if fill_x:
        x = np.full(nbins, np.nan)
    edge = lo                                  # the reference accumulates the edge by repeated addition: do the same
    for k in range(nbins):
        inside = (signal >= edge) & (signal <= edge + step)       # both ends inclusive, as in the reference
        edge += step
        d = noise[inside]
        if len(d) >= min_len:
            weights[k] = len(d)
            y[k] = average(d)
            if fill_x:
                x[k] = edge - 0.5 * step
    return x, y, weights, signal


def _validI(x, y, weights):
    """bins worth fitting: finite and better populated than the median bin.  (The reference also computes a gradient
    based outlier mask but assigns it through ``i[i][...] = False``, a copy — it never takes effect; nothing to mirror.)"""
    return np.logical_and(np.isfinite(y), weights > np.median(weights))


def _evaluate(x, y, weights):
    """square-root fit through the valid bins -> (fitParams or None, callable, valid-bin mask)"""
    from scipy.optimize import curve_fit
    i = _validI(x, y, weights)
    xx, yy = x[i], y[i]
    try:
        popt, _ = curve_fit(function, xx, yy, check_finite=False)
        min_y = function(xx[0], *popt)
        params = np.insert(popt, 0, min_y)
        return params, FittedNLF(params=params), i
    except RuntimeError:
        print("couldn't fit noise function with filtered indices, use polynomial fit instead")
        return None, smooth(xx, yy, weights[i]), i


def smooth(x, y, weights):
    """fallback when the square-root model does not fit: weighted 2nd-degree polynomial, clamped to the fitted range;
    a constant if even that fails"""
    p = np.polyfit(x, y, w=weights, deg=2)
    if np.any(np.isnan(p)):
        return FittedNLF(const=np.average(y, weights=weights))
    return FittedNLF(poly=p, xrange=(x[0], x[-1]))


def oneImageNLF(img, img2=None, signal=None):
    """estimate the noise level function from one image (or two of the same scene) -> (callable, signal image)"""
    x, y, weights, signal = calcNLF(img, img2, signal)
    _, fn, _ = _evaluate(x, y, weights)
    return fn, signal
