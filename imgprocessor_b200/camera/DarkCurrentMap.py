"""Dark-current map producers on the GPU — mirror of imgProcessor.camera.DarkCurrentMap
(/root/reference/imgProcessor/camera/DarkCurrentMap.py):

  * DarkCurrentMap / averageSameExpTimes (:16-58): single-time-effect-free average of background images (nStd = 3) —
    kernel K4, the noise level function estimated from the images as the reference does when none is given;
  * getLinearityFunction (:61-80): per-pixel line image(t) = offset + ascent * t over the exposure series — kernel K5
    (csrc/k5_producers.cu).  The regression restates fancytools.linRegressUsingMasked2dArrays, which is ABSENT from the
    reference tree (un-vendored dependency): parity for that ingredient is unpinned (oracle/producers.py);
  * sortForSameExpTime / getDarkCurrentAverages / getDarkCurrentFunction (:83-124): host bookkeeping.
The (offset, ascent) pair is what CameraCalibration.addDarkCurrent(slope, intercept) stores and K1 evaluates per pixel.
"""
from collections import OrderedDict

import numpy as np

from .. import engine as _engine
from ..imgIO import imread
from . import NoiseLevelFunction as _nlf


class DarkCurrentMap(object):
    """average of background images with single time effects removed (the images are collected and reduced on the GPU
    when the map is asked for)"""

    def __init__(self, twoImages, noise_level_function=None, calcVariance=False, **kwargs):
        assert len(twoImages) > 1, 'need at least 2 images'
        if calcVariance:
            raise NotImplementedError('variance tracking (fancytools.MaskedMovingAverage(calcVariance=True)) is not on the GPU path')
        self._images = [np.asarray(imread(i)) for i in twoImages]
        self._nlf = noise_level_function
        self._map = None

    def addImg(self, img, raiseIfConvergence=False):
        if raiseIfConvergence:
            raise NotImplementedError('convergence check needs the variance map')
        self._images.append(np.asarray(imread(img)))
        self._map = None

    def map(self):
        if self._map is None:
            frames = self._images
            dt = np.result_type(*[f.dtype for f in frames])
            if dt.type not in (np.uint8, np.uint16, np.float32, np.float64):
                dt = np.dtype(np.float64)
            stack = np.stack([np.ascontiguousarray(f, dtype=dt) for f in frames])
            nlf = self._nlf
            first = np.min((stack[0].astype(np.float64), stack[1]), axis=0)
            if nlf is None:
                nlf = _nlf.oneImageNLF(first)[0]          # SingleTimeEffectDetection.py:43-45
            eng = _engine.get_engine(*stack.shape[1:])
            tt = _engine.torch()
            dev = tt.from_numpy(stack).to(eng.device)
            params = getattr(nlf, 'params', None) or getattr(nlf, 'coeff', None)
            if params is not None:
                avg = eng.ste_average(dev, params, 3.0)
            else:
                avg = eng.ste_average(dev, threshold=np.array(np.broadcast_to(np.asarray(nlf(first), np.float64) * 3.0, first.shape)))
            self._map = _engine.to_numpy(avg)
        return self._map


def averageSameExpTimes(imgs_path):
    """average background images taken with the same exposure time"""
    d = DarkCurrentMap([np.asarray(imread(i), dtype=np.float64) for i in imgs_path[:2]])
    for i in imgs_path[2:]:
        d.addImg(i)
    return d.map()


def getLinearityFunction(expTimes, imgs, mxIntensity=65535, min_ascent=0.001):
    """offset, ascent, error of image(expTime) = offset + ascent * expTime per pixel"""
    imgs = np.ascontiguousarray(imgs)
    if imgs.dtype.type not in (np.uint8, np.uint16, np.float32, np.float64):
        imgs = imgs.astype(np.float64)
    tt = _engine.torch()
    eng = _engine.get_engine(*imgs.shape[1:])
    offset, ascent, error = eng.linear_fit(tt.from_numpy(imgs).to(eng.device), expTimes, mxIntensity, min_ascent)
    return _engine.to_numpy(offset), _engine.to_numpy(ascent), _engine.to_numpy(error)


def sortForSameExpTime(expTimes, img_paths):
    """exposure times (sorted) and, for each, the images taken with it"""
    d = {}
    for e, i in zip(expTimes, img_paths):
        d.setdefault(e, []).append(i)
    d = OrderedDict(sorted(d.items()))
    return list(d.keys()), list(d.values())


def getDarkCurrentAverages(exposuretimes, imgs):
    """exposure times and one (averaged) image per exposure time"""
    x, groups = sortForSameExpTime(exposuretimes, imgs)
    s0, s1 = imgs[0].shape
    out = np.empty((len(x), s0, s1), dtype=imgs[0].dtype)
    for o, g in zip(out, groups):
        o[:] = g[0] if len(g) == 1 else averageSameExpTimes(g)
    return x, out


def getDarkCurrentFunction(exposuretimes, imgs, **kwargs):
    """dark current as a function of the exposure time: offset, ascent, rmse maps"""
    exposuretimes, imgs = getDarkCurrentAverages(exposuretimes, imgs)
    return getLinearityFunction(exposuretimes, imgs, **kwargs)
