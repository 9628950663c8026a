"""flatFieldFromCloseDistance on the GPU — mirror of
imgProcessor.camera.flatField.flatFieldFromCloseDistance.flatFieldFromCloseDistance
(/root/reference/imgProcessor/camera/flatField/flatFieldFromCloseDistance.py:16-38, "Method A"):

    img = imgAverage(imgs)                         transform/imgAverage.py:7-22   (float64, summed in order)
    bg  = getBackground2(bg_imgs, img)             utils/getBackground2.py:5-12   (average of bg_imgs, or a number)
    img -= bg
    img = toGray(img)                              transformations.py:126-135     (colour frames: luminance weights)
    mx  = max of the 3x3 median filter of img[::10, ::10]
    img /= mx

Kernel K5 (csrc/k5_producers.cu) does the average / background / luminance in one streaming pass and the final scaling,
K1 the 3x3 median of the subsampled image.  As shipped the reference only works for COLOUR frames (toGray averages the
last axis with three weights); gray frames make it raise, and so does this mirror.  ``bg_imgs=None`` needs
imgSignal.scaleSignalCutParams -> fancytools.findXAt, which is absent from the reference tree: not available.
"""
import numpy as np

from ... import engine as _engine
from ...imgIO import imread


def _stack(images):
    frames = [np.asarray(imread(i)) for i in images]
    if any(f.shape != frames[0].shape for f in frames):
        raise ValueError('images of different shapes')
    dt = np.result_type(*[f.dtype for f in frames])
    if dt.type not in (np.uint8, np.uint16, np.float32, np.float64):
        dt = np.dtype(np.float64)
    return np.stack([np.ascontiguousarray(f, dtype=dt) for f in frames])


def flatFieldFromCloseDistance(imgs, bg_imgs=None):
    """average of colour images of a homogeneous device right in front of the lens, background removed, luminance, scaled by
    the maximum of the 3x3 median of every tenth pixel -> float64 [H,W]"""
    if bg_imgs is None:
        raise NotImplementedError('bg_imgs=None estimates the background through imgSignal.scaleSignalCutParams -> '
                                  'fancytools.findXAt, which the reference tree does not contain; pass background images or a number')
    stack = _stack(imgs)
    if stack.ndim != 4 or stack.shape[-1] != 3:
        # the reference: np.average(img, axis=-1, weights=(0.299, 0.587, 0.114)) fails unless the last axis has 3 entries
        raise ValueError('Length of weights not compatible with specified axis.')
    tt = _engine.torch()
    H, W = stack.shape[1:3]
    eng = _engine.get_engine(H, W)
    dev = tt.from_numpy(stack).to(eng.device)
    if type(bg_imgs) in (int, float):
        gray = eng.stack_mean(dev, minus=bg_imgs, gray3=True)
    else:
        bg = eng.stack_mean(tt.from_numpy(_stack(bg_imgs)).to(eng.device))
        gray = eng.stack_mean(dev, minus=bg, gray3=True)
    small = eng.subsample(gray, 10, 10)
    h, w = small.shape
    med = _engine.get_engine(h, w).median3x3(_engine.to_numpy(small))
    eng.scale_(gray, float(med.max()))
    return _engine.to_numpy(gray)
