"""Host logic of the noise-level-function estimate (imgprocessor_b200/camera/NoiseLevelFunction.py) against outputs of the
unmodified reference (tests/golden/ste_nlf.npz) — CPU only: the median image is taken from the golden file, everything
after it (bins, weights, curve_fit, polynomial fallback) is what this module computes."""
import numpy as np

from conftest import load_golden
from imgprocessor_b200.camera import NoiseLevelFunction as nlfmod


def test_binning_and_fit_match_the_reference():
    g = load_golden('ste_nlf')
    x, y, w, signal = nlfmod.calcNLF(g['avg0'], signal=g['nlf_signal'])
    filled = g['nlf_w'] > 0
    assert signal is not None and np.array_equal(w, g['nlf_w'])
    assert np.array_equal(x[filled], g['nlf_x'][filled]) and np.array_equal(y[filled], g['nlf_y'][filled])
    assert np.isnan(y[~filled]).all()
    params, fn, valid = nlfmod._evaluate(x, y, w)
    assert np.array_equal(valid, g['nlf_valid']) and np.array_equal(params, g['nlf_params'])
    assert fn.params == tuple(g['nlf_params']) and np.array_equal(fn(g['nlf_curve_x']), g['nlf_curve'])
    x2, y2, w2, _ = nlfmod.calcNLF(g['frames'][0], g['frames'][1], signal=g['nlf2_signal'])
    f2 = g['nlf2_w'] > 0
    assert np.array_equal(w2, g['nlf2_w']) and np.array_equal(y2[f2], g['nlf2_y'][f2]) and np.array_equal(x2[f2], g['nlf2_x'][f2])


def test_polynomial_fallback_and_bounded_function():
    g = load_golden('ste_nlf')
    sm = nlfmod.smooth(g['smooth_x'], g['smooth_y'], g['smooth_w'])
    assert sm.params is None and np.array_equal(sm(g['smooth_eval_x']), g['smooth_eval'])
    s = load_golden('ste')
    assert np.array_equal(nlfmod.boundedFunction(s['bf_x'], *s['nlf']), s['bf_y'], equal_nan=True)
    const = nlfmod.smooth(np.array([1.0, 1.0, 1.0]), np.array([2.0, 4.0, 6.0]), np.array([1.0, 1.0, 2.0]))
    assert np.isfinite(const(np.array([0.0, 5.0]))).all()
