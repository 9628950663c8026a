"""medianThreshold on the B200: drop-in for imgProcessor.filters.medianThreshold.medianThreshold
(filters/medianThreshold.py:7-30).  The NxN median (scipy.ndimage 'reflect' border), the float64
relative-deviation predicate and the masked replacement all run inside kernel K1
(csrc/k1_pointwise_median.cu); there is no CPU fallback."""
import numpy as np

from .. import engine as _engine

_NP_OK = (np.uint8, np.uint16, np.float32, np.float64)


def medianThreshold(img, threshold=0.1, size=3, condition='>', copy=True):
    """Set every pixel of ``img`` whose relative deviation from the size x size median,
    abs((img - median) / median), exceeds (``condition='>'``) or stays below (``'<'``) ``threshold``
    to that median.  Returns ``(img, indices)``; ``indices`` is None and ``img`` is returned
    untouched when ``threshold <= 0``.  ``copy=False`` modifies and returns the very same array.

    ``img`` may be a 2-D numpy array (uint8, uint16, float32, float64 — the dtype is kept) or a CUDA
    torch tensor of those dtypes (the result then stays on the device)."""
    if not threshold > 0:
        return img, None
    if size not in (3, 5):
        raise ValueError('size must be 3 or 5 (got %r)' % (size,))
    if condition not in ('>', '<'):
        condition = '<'          # the reference treats everything that is not '>' as '<'
    tt = _engine.torch()
    if isinstance(img, tt.Tensor):
        if img.dim() != 2:
            raise ValueError('medianThreshold works on single 2-D frames')
        eng = _engine.get_engine(img.shape[0], img.shape[1], img.device)
        out, ind = eng.pointwise_median(img, threshold, size, condition, flags=0, out_dtype=img.dtype, want_mask=True)
        ind = ind.bool()
        if not copy:
            img.copy_(out)
            out = img
        return out, ind
    if not isinstance(img, np.ndarray) or img.ndim != 2:
        raise ValueError('medianThreshold works on single 2-D frames')
    if img.dtype.type not in _NP_OK:
        raise TypeError('unsupported image dtype %s (uint8, uint16, float32, float64)' % img.dtype)
    eng = _engine.get_engine(img.shape[0], img.shape[1])
    dev = tt.from_numpy(np.ascontiguousarray(img)).to(eng.device)
    out, ind = eng.pointwise_median(dev, threshold, size, condition, flags=0, out_dtype=dev.dtype, want_mask=True)
    res = _engine.to_numpy(out)
    indices = ind.cpu().numpy().astype(bool)
    if copy:
        return res, indices
    img[...] = res
    return img, indices
