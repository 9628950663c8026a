"""The per-pixel arithmetic the CUDA kernels are built from (imgcorr_core.cuh), compiled for the
host and checked bit for bit against the oracle models — no GPU needed."""
import numpy as np
import pytest
from hypothesis import given, settings, strategies as st

import emul
from conftest import load_golden
from imgprocessor_b200 import synth
from oracle import models, refpath


def _case(H, W, seed, dtype=np.uint16):
    raw = synth.scene(H, W, seed, dtype)
    dark = synth.dark_map(H, W, seed)
    flat = synth.flat_map(H, W, seed, p_zero=5e-3)
    return raw, dark, flat


@pytest.mark.parametrize('shape', [(1, 1), (2, 3), (5, 4), (33, 130), (96, 128), (70, 257)])
@pytest.mark.parametrize('dtype', [np.uint8, np.uint16, np.float32])
@pytest.mark.parametrize('ksize', [3, 5])
def test_k1_chain_bit_exact(shape, dtype, ksize):
    H, W = shape
    raw, dark, flat = _case(H, W, 3, dtype)
    got, mask = emul.k1(raw, dark, flat, 0.1, ksize)
    x = models.pointwise_model(raw, dark, flat, nan_to_num=True)
    want, wmask = models.median_threshold_model(x, 0.1, ksize)
    assert np.array_equal(got, want) and np.array_equal(mask, wmask)


def test_k1_pointwise_zero_ulp_vs_float64_reference():
    raw, dark, flat = _case(64, 96, 9)
    flat[3, 3] = np.float32(1e-42)
    flat[4, 4] = np.float32(-2.0)
    got, _ = emul.k1(raw, dark, flat, 0.0, 0, flags=3)
    x = refpath.to_float_image(raw)
    refpath.correct_dark_current(x, dark)
    refpath.correct_vignetting(x, flat)
    with np.errstate(over='ignore'):
        assert np.array_equal(got, x.astype(np.float32))
    assert np.isinf(got[3, 3])
    got, _ = emul.k1(raw, dark, flat, 0.0, 0, flags=7)
    assert got[3, 3] == np.finfo(np.float32).max


def test_k1_legacy_linear_dark():
    g = load_golden('correct_u16_legacy_dark')
    offs, asc = g['offs'].astype(np.float32), g['ascent'].astype(np.float32)
    got, _ = emul.k1(g['raw'], offs, None, 0.0, 0, flags=1, ascent=asc, exposure=float(g['exposure_time']))
    want = models.pointwise_model(g['raw'], offs, None, False, dark_ascent=asc, exposure_time=float(g['exposure_time']))
    assert np.array_equal(got, want)


@pytest.mark.parametrize('name', ['median_u16_s3_gt', 'median_u16_s3_lt', 'median_u16_s5_gt', 'median_u16_s5_lt',
                                  'median_f32_s3_gt', 'median_f32_s3_lt', 'median_f32_s5_gt', 'median_f32_s5_lt'])
def test_k1_direct_median_vs_reference_golden(name):
    g = load_golden(name)
    size = 3 if '_s3_' in name else 5
    cond = '>' if name.endswith('gt') else '<'
    got, mask = emul.k1(g['img'], None, None, 0.1, size, cond, flags=0, out_dtype=g['img'].dtype)
    assert got.dtype == g['out'].dtype
    assert np.array_equal(got, g['out']) and np.array_equal(mask, g['ind'])


def test_k1_zero_medians_golden():
    g = load_golden('median_f32_zeros')
    got, mask = emul.k1(g['img'], None, None, float(g['threshold']), 3, flags=0)
    assert np.array_equal(got, g['out']) and np.array_equal(mask, g['ind'])


def test_k1_float64_frames():
    rng = np.random.default_rng(4)
    img = rng.normal(1000, 300, (40, 50))
    img[rng.random(img.shape) < 0.01] *= 5
    for size in (3, 5):
        got, mask = emul.k1(img, None, None, 0.2, size, flags=0)
        want, wmask = models.median_threshold_model(img, 0.2, size)
        assert got.dtype == np.float64 and np.array_equal(got, want) and np.array_equal(mask, wmask)


@settings(max_examples=60, deadline=None)
@given(st.integers(0, 2 ** 31 - 1), st.sampled_from([3, 5]), st.floats(1e-3, 2.0), st.sampled_from(['>', '<']))
def test_k1_predicate_guard_band_random(seed, ksize, thr, cond):
    """random float32 frames incl. values engineered to sit on the threshold: the float32 fast
    path + float64 fallback must reproduce the float64 predicate exactly."""
    rng = np.random.default_rng(seed)
    H, W = 12, 16
    img = (rng.random((H, W)) * 100 + 1).astype(np.float32)
    med = models.median_filter_reflect(img, ksize)
    # push a few pixels to |x-b|/b == thr up to float32 rounding, from both sides
    for _ in range(6):
        y, x = rng.integers(0, H), rng.integers(0, W)
        b = np.float64(med[y, x])
        img[y, x] = np.float32(b * (1 + thr * rng.choice([-1, 1])) * (1 + rng.choice([-1, 0, 1]) * 6e-8))
    got, mask = emul.k1(img, None, None, thr, ksize, cond, flags=0)
    want, wmask = models.median_threshold_model(img, thr, ksize, cond)
    assert np.array_equal(mask, wmask) and np.array_equal(got, want)


def test_k1_extreme_values():
    img = np.array([[3e38, -3e38, 1e-40, 0, 5], [np.inf, -np.inf, np.nan, 1, 2], [0, 0, 0, 0, 0]], np.float32)
    for thr in (0.1, 5.0):
        got, mask = emul.k1(img, None, None, thr, 3, flags=4)
        x = models.nan_to_num_f32(img)
        want, wmask = models.median_threshold_model(x, thr, 3)
        assert np.array_equal(mask, wmask) and np.array_equal(got, want)


@pytest.mark.parametrize('tag', ['moderate', 'strong', 'realistic'])
def test_maps_vs_reference_golden(tag):
    g = load_golden('maps_' + tag)
    H, W = (int(v) for v in g['shape'])
    mx, my = emul.maps(g['K'], g['dist'], g['P'], H, W)
    for a, b in zip(models.fixed_point_coords(mx, my), models.fixed_point_coords(g['mapx'], g['mapy'])):
        assert np.array_equal(a, b)
    assert np.abs(mx.astype(np.float64) - g['mapx']).max() < 2e-4
    assert (mx != g['mapx']).mean() < 1e-3 and (my != g['mapy']).mean() < 1e-3


def test_maps_full_size_fixed_point_agreement():
    """4096x3000 moderate lens (BASELINE configs[1..2]): every 5-bit coordinate equals OpenCV's."""
    H, W = 3000, 4096
    p = synth.lens_moderate(H, W)
    K, d = synth.camera_matrix(p), synth.dist_coeffs(p)
    mapx, mapy, P, roi = refpath.undistort_rectify_map(K, d, W, H)
    mx, my = emul.maps(K, d, P, H, W)
    for a, b in zip(models.fixed_point_coords(mx, my), models.fixed_point_coords(mapx, mapy)):
        assert np.array_equal(a, b)
    assert (mx != mapx).mean() < 1e-4 and (my != mapy).mean() < 1e-4


@pytest.mark.parametrize('dtype', [np.uint8, np.uint16, np.float32, np.float64])
def test_remap_vs_model_and_golden(dtype):
    tag = {np.uint8: 'u8', np.uint16: 'u16', np.float32: 'f32', np.float64: 'f64'}[dtype]
    g1 = load_golden('lens_f32_keep1')
    for keep in (0, 1):
        g = load_golden('lens_%s_keep%d' % (tag, keep))
        H, W = g['img'].shape
        x, y, w, h = (int(v) for v in g['roi'])
        win = None if keep else (x, y, w, h)
        got = emul.remap(g['img'], g1['mapx'], g1['mapy'], float(g['border']), win)
        assert got.dtype == g['out'].dtype and np.array_equal(got, g['out'])


def test_remap_weird_coordinates():
    rng = np.random.default_rng(3)
    H, W = 40, 50
    src = (rng.random((H, W)) * 1000).astype(np.float32)
    mapx = np.arange(W, dtype=np.float32)[None, :] + rng.normal(0, 4, (H, W)).astype(np.float32)
    mapy = np.arange(H, dtype=np.float32)[:, None] + rng.normal(0, 4, (H, W)).astype(np.float32)
    mapx[0, :6] = [-1.0, -0.5, W - 1, W - 0.5, W, 1e9]
    mapy[1, :4] = [-1.0, H - 1, H, -1e9]
    mapx[2, :3] = [np.nan, np.inf, -np.inf]
    for border in (0.0, 0.1234567):
        assert np.array_equal(emul.remap(src, mapx, mapy, border), models.remap_model(src, mapx, mapy, border))
    w = emul.remap(src, mapx, mapy, 0.0, widen=True)
    assert w.dtype == np.float64 and np.array_equal(w, models.remap_model(src, mapx, mapy, 0).astype(np.float64))


def test_full_chain_vs_reference_golden():
    """emulated K1 -> K2 against the unmodified float64 reference output."""
    g = load_golden('correct_u16_keep1')
    x, _ = emul.k1(g['raw'], g['dark'], g['flat'], 0.1, 3)
    _, _, P, _ = refpath.undistort_rectify_map(g['K'], g['dist'], 128, 96)
    mx, my = emul.maps(g['K'], g['dist'], P, 96, 128)
    out = emul.remap(x, mx, my, 0.0)
    ok = np.abs(g['out']) < 1e6
    assert np.abs(out[ok] - g['out'][ok]).max() / 65535.0 < 1e-5
    want, _ = models.correct_chain_f32(g['raw'], g['dark'], g['flat'], 0.1, 3, mapxy=(mx, my))
    assert np.array_equal(out, want)


def test_order_preserving_keys():
    """the 5x5 streaming kernel sorts integer keys instead of floats: the mapping must be a bijection that preserves
    the order PTX min / max use (-0.0 below +0.0), over denormals, infinities and the whole finite range"""
    rng = np.random.default_rng(12)
    x = np.concatenate([rng.normal(0, 1e3, 5000), rng.normal(0, 1e-40, 500), [0.0, -0.0, np.inf, -np.inf, 3.4e38, -3.4e38,
                        1e-45, -1e-45, 1.0, -1.0]]).astype(np.float32)
    k, back = emul.keys(x)
    assert np.array_equal(back.view(np.uint32), x.view(np.uint32))              # exact round trip, sign of zero included
    order = np.argsort(k, kind='stable')
    xs = x[order]
    assert np.all(xs[:-1] <= xs[1:])                                            # same order as the floats
    kz, _ = emul.keys(np.array([-0.0, 0.0], np.float32))
    assert kz[0] < kz[1]
