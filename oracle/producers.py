"""Calibration-map producers, restated for the CPU.

TEST INFRASTRUCTURE ONLY (see oracle/__init__.py).

Reference (paths relative to /root/reference/imgProcessor/):
  * flatFieldFromCloseDistance          camera/flatField/flatFieldFromCloseDistance.py:16-38
      imgAverage transform/imgAverage.py:7-22, getBackground2 utils/getBackground2.py:5-12, toGray transformations.py:126-135
  * getLinearityFunction                camera/DarkCurrentMap.py:61-80
  * averageSameExpTimes / DarkCurrentMap  camera/DarkCurrentMap.py:16-58  (-> oracle/ste.py with nStd = 3)

Pinned by tests/golden/producers.npz (outputs of the unmodified reference) except for ONE ingredient:
``fancytools.math.linRegressUsingMasked2dArrays`` (getLinearityFunction's regression) lives in the un-vendored, un-pinned
fancytools package, absent here — PARITY UNPINNED.  Its published algorithm, ordinary least squares per pixel over the
samples that are not masked plus the RMSE of the residuals, is restated in ``lin_regress_masked``.
"""
import numpy as np


def img_average(images):
    out = np.array(images[0], dtype=np.float64)
    for i in images[1:]:
        out += np.asarray(i, dtype=np.float64)
    out /= len(images)
    return out


def flat_field_from_close_distance(imgs, bg_imgs):
    from scipy.ndimage import median_filter
    img = img_average(imgs)
    bg = bg_imgs if type(bg_imgs) in (int, float) else img_average(bg_imgs)
    img -= bg
    img = np.average(img, axis=-1, weights=(0.299, 0.587, 0.114)).astype(img.dtype)
    mx = median_filter(img[::10, ::10], 3).max()
    img /= mx
    return img


def lin_regress_masked(x, arrays, bad_mask):
    """restatement of fancytools.math.linRegressUsingMasked2dArrays(xVals, arrays, badMask): see module docstring.
    Sums run over the exposure index in order; returns ascent, offset, rmse."""
    x = np.asarray(x, np.float64)
    y = np.asarray(arrays, np.float64)
    good = ~np.asarray(bad_mask, bool)
    sx = np.zeros(y.shape[1:]); sy = np.zeros(y.shape[1:]); sxy = np.zeros(y.shape[1:]); sxx = np.zeros(y.shape[1:])
    m = np.zeros(y.shape[1:])
    for k in range(len(x)):
        g = good[k]
        sx = np.where(g, sx + x[k], sx)
        sy = np.where(g, sy + y[k], sy)
        sxy = np.where(g, sxy + x[k] * y[k], sxy)
        sxx = np.where(g, sxx + x[k] * x[k], sxx)
        m = np.where(g, m + 1, m)
    with np.errstate(divide='ignore', invalid='ignore'):
        ascent = (m * sxy - sx * sy) / (m * sxx - sx * sx)
        offset = (sy - ascent * sx) / m
        e = np.zeros(y.shape[1:])
        for k in range(len(x)):
            r = y[k] - (offset + ascent * x[k])
            e = np.where(good[k], e + r * r, e)
        rmse = np.sqrt(e / m)
    return ascent, offset, rmse


def get_linearity_function(exp_times, imgs, mx_intensity=65535, min_ascent=0.001):
    imgs = np.asarray(imgs)
    ascent, offset, error = lin_regress_masked(exp_times, imgs, imgs > mx_intensity)
    ascent[np.isnan(ascent)] = 0
    if min_ascent > 0:
        i = ascent < min_ascent
        offset[i] += (0.5 * (np.min(exp_times) + np.max(exp_times))) * ascent[i]
        ascent[i] = 0
    return offset, ascent, error
