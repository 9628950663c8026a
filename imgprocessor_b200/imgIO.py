"""Input normalisation used by the correct() path: mirrors the subset of
imgProcessor.imgIO.imread (imgIO.py:39-73) the path exercises — array pass-through, callables,
file paths via cv2.imread, optional dtype cast (``img.astype(dtype)`` for float targets, :17-22)."""
import numpy as np


def imread(img, color=None, dtype=None):
    if callable(img):
        img = img()
    elif isinstance(img, str):
        import cv2
        flag = {'gray': cv2.IMREAD_GRAYSCALE, 'all': cv2.IMREAD_COLOR, None: cv2.IMREAD_ANYCOLOR}[color]
        if dtype in (None, 'noUint') or np.dtype(dtype) != np.uint8:
            flag |= cv2.IMREAD_ANYDEPTH
        path = img
        img = cv2.imread(path, flag)
        if img is None:
            raise IOError("image '%s' is not existing" % path)
    elif color == 'gray' and getattr(img, 'ndim', 2) == 3:
        # luminance of an RGB(A) array, as imgProcessor.transformations.toGray does for float input
        img = np.average(img[..., :3], axis=-1, weights=(0.299, 0.587, 0.114))
    if dtype is not None and isinstance(img, np.ndarray):
        dt = np.dtype(float if dtype == 'float' else dtype) if dtype != 'noUint' else None
        if dt is None:
            if img.dtype.kind == 'u':
                img = img.astype(np.int32 if img.dtype.itemsize < 4 else np.int64)
        elif dt.kind in 'ui':
            if img.dtype != dt:
                if img.dtype.kind in 'ui' and np.iinfo(dt).max >= img.max(initial=0) and img.min(initial=0) >= 0:
                    img = img.astype(dt)
                else:
                    raise NotImplementedError('range-scaling conversion to %s is outside the correct() path' % dt)
        else:
            img = img.astype(dt)
    return img
