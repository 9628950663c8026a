"""Pin both oracle layers against outputs of the UNMODIFIED reference
(tests/golden/*.npz, made by tests/golden/make_golden.py in the build container)."""
import numpy as np
import pytest

from oracle import models, refpath
from conftest import load_golden, ulp_diff_f32


@pytest.mark.parametrize('keep', [1, 0])
def test_refpath_full_chain_bit_exact(keep):
    g = load_golden('correct_u16_keep%d' % keep)
    out = refpath.correct(g['raw'], g['dark'], g['flat'], (g['K'], g['dist']), threshold=0.1,
                          keep_size=bool(keep))
    assert out.dtype == np.float64 and out.shape == g['out'].shape
    assert np.array_equal(out, g['out'])


def test_refpath_last_img_is_pre_median():
    g = load_golden('correct_u16_keep1')
    x = refpath.to_float_image(g['raw'])
    refpath.correct_dark_current(x, g['dark'])
    refpath.correct_vignetting(x, g['flat'])
    assert np.array_equal(x, g['last_img'], equal_nan=True)
    assert np.array_equal(g['bg'], g['dark'])


def test_refpath_thr0_and_partial_calibrations():
    g = load_golden('correct_f32_thr0')
    out = refpath.correct(g['raw'], g['dark'], g['flat'], (g['K'], g['dist']), threshold=0)
    assert np.array_equal(out, g['out'], equal_nan=True)
    g = load_golden('correct_f32_flat_only')
    assert np.array_equal(refpath.correct(g['raw'], None, g['flat'], None, threshold=0.25), g['out'])
    g = load_golden('correct_u16_nothing')
    assert np.array_equal(refpath.correct(g['raw'], threshold=0.1), g['out'])


def test_refpath_legacy_dark_tuple():
    g = load_golden('correct_u16_legacy_dark')
    bg = refpath.calc_dark_current((None, '', (g['offs'].copy(), g['ascent'].copy()), None),
                                   float(g['exposure_time']))
    assert np.array_equal(bg, g['bg'])
    assert np.array_equal(refpath.correct(g['raw'], bg, threshold=0.1), g['out'])


@pytest.mark.parametrize('tag', ['u16', 'f32'])
@pytest.mark.parametrize('size', [3, 5])
@pytest.mark.parametrize('cond', ['gt', 'lt'])
def test_median_threshold_both_layers(tag, size, cond):
    g = load_golden('median_%s_s%d_%s' % (tag, size, cond))
    c = '>' if cond == 'gt' else '<'
    o, ind = refpath.median_threshold(g['img'], 0.1, size, c, copy=True)
    assert o.dtype == g['out'].dtype and np.array_equal(o, g['out']) and np.array_equal(ind, g['ind'])
    o, ind = models.median_threshold_model(g['img'], 0.1, size, c)
    assert o.dtype == g['out'].dtype and np.array_equal(o, g['out']) and np.array_equal(ind, g['ind'])


def test_median_threshold_zero_medians():
    g = load_golden('median_f32_zeros')
    o, ind = models.median_threshold_model(g['img'], float(g['threshold']), 3)
    assert np.array_equal(o, g['out']) and np.array_equal(ind, g['ind'])
    # blur==0 & img==0 -> nan -> False ; blur==0 & img!=0 -> inf -> True (pixel becomes 0)
    assert not ind[25, 30] and ind[20, 20] and o[20, 20] == 0


@pytest.mark.parametrize('tag', ['u8', 'u16', 'f32', 'f64'])
@pytest.mark.parametrize('keep', [0, 1])
def test_lens_correct_both_layers(tag, keep):
    g = load_golden('lens_%s_keep%d' % (tag, keep))
    out = refpath.lens_correct(g['img'], g['K'], g['dist'], keep_size=bool(keep), border_value=float(g['border']))
    assert out.dtype == g['out'].dtype and np.array_equal(out, g['out'])
    h, w = g['img'].shape
    mapx, mapy, P, roi = refpath.undistort_rectify_map(g['K'], g['dist'], w, h)
    assert tuple(roi) == tuple(g['roi'])
    m = models.remap_model(g['img'], mapx, mapy, float(g['border']))
    if not keep:
        x, y, ww, hh = roi
        m = m[y:y + hh, x:x + ww]
    assert m.dtype == g['out'].dtype and np.array_equal(m, g['out'])


@pytest.mark.parametrize('tag', ['moderate', 'strong', 'realistic'])
def test_map_model_vs_reference_maps(tag):
    g = load_golden('maps_' + tag)
    h, w = (int(v) for v in g['shape'])
    mx, my = models.undistort_map_model(g['K'], g['dist'], g['P'], w, h)
    # float64 evaluation order differs from OpenCV's incremental/SIMD loop: allow 1 float32 ulp
    # on a vanishing fraction of entries, but the 5-bit fixed-point coordinates must all agree
    for a, b in ((mx, g['mapx']), (my, g['mapy'])):
        d = ulp_diff_f32(a, b)
        tiny = np.abs(b) < 1e-3          # alpha=1 maps the corner pixel to ~0: cancellation, ulps meaningless
        assert d[~tiny].max() <= 1 and (d[~tiny] > 0).mean() < 1e-3
        assert np.abs(a[tiny].astype(np.float64) - b[tiny]).max(initial=0) < 1e-9
    fa = models.fixed_point_coords(mx, my)
    fb = models.fixed_point_coords(g['mapx'], g['mapy'])
    for a, b in zip(fa, fb):
        assert np.array_equal(a, b)


def test_f32_chain_close_to_f64_reference():
    """the float32 staged chain (what the GPU computes) against the float64 reference:
    same threshold decisions, deviation at float32 rounding level."""
    g = load_golden('correct_u16_keep1')
    out32, mask = models.correct_chain_f32(g['raw'], g['dark'], g['flat'], 0.1, 3, None)
    ref64 = refpath.correct(g['raw'], g['dark'], g['flat'], None, 0.1)
    big = np.abs(ref64) > 3e38          # denormal-flat pixels overflow float32: nan_to_num clamps them
    assert big.sum() <= 4
    err = ulp_diff_f32(out32[~big], ref64[~big].astype(np.float32))
    assert err.max() <= 1
    # and through the lens: <= 1e-3 of full scale end to end (north star), in practice ~1e-7
    mapx, mapy, P, _ = refpath.undistort_rectify_map(g['K'], g['dist'], 128, 96)
    out32, _ = models.correct_chain_f32(g['raw'], g['dark'], g['flat'], 0.1, 3, (g['K'], g['dist'], P))
    ok = np.abs(g['out']) < 1e6
    assert np.abs(out32[ok] - g['out'][ok]).max() / 65535.0 < 1e-5


def test_date_selection_restated():
    g = load_golden('correct_u16_dates')
    raw = g['raw'].astype(np.float64)
    # the list is kept newest-first (CameraCalibration.py:29-34); _getFromDate (:37-49) returns the
    # entry just BEFORE the first one older than `date`, i.e. the oldest calibration that is not
    # older than the requested date, and the newest one when every entry is older / date is
    # None or unparsable.
    assert np.array_equal(g['out_none'], raw - g['d3'])
    assert np.array_equal(g['out_bad'], raw - g['d3'])
    assert np.array_equal(g['out_new'], raw - g['d3'])
    assert np.array_equal(g['out_mid'], raw - g['d3'])
    assert np.array_equal(g['out_old'], raw - g['d1'])
