"""The pure-numpy models of scipy's reflect rank filter and OpenCV's map + remap
arithmetic against the live libraries (scipy 1.18.1 / OpenCV 4.13.0 in the image)."""
import numpy as np
import pytest

from oracle import models, refpath
from imgprocessor_b200 import synth

cv2 = pytest.importorskip('cv2')
ndi = pytest.importorskip('scipy.ndimage')


@pytest.mark.parametrize('shape', [(1, 1), (1, 7), (2, 2), (3, 5), (17, 31), (64, 80)])
@pytest.mark.parametrize('size', [3, 5])
@pytest.mark.parametrize('dtype', [np.uint16, np.float32, np.float64])
def test_median_reflect_vs_scipy(shape, size, dtype):
    rng = np.random.default_rng(hash((shape, size)) % 2 ** 31)
    img = (rng.random(shape) * 5000 - 1000).astype(dtype) if dtype != np.uint16 else \
        rng.integers(0, 65536, shape).astype(dtype)
    assert np.array_equal(models.median_filter_reflect(img, size), ndi.median_filter(img, size=size))


def test_median_ties_and_signed_zero():
    img = np.zeros((9, 9), np.float32)
    img[::2] = -0.0
    img[4, 4] = 3
    a = models.median_filter_reflect(img, 3)
    b = ndi.median_filter(img, size=3)
    assert np.array_equal(a, b)          # value equality; sign of zero is not observable downstream


def _rand_maps(rng, H, W, spread):
    mapx = np.arange(W, dtype=np.float32)[None, :] + rng.normal(0, spread, (H, W)).astype(np.float32)
    mapy = np.arange(H, dtype=np.float32)[:, None] + rng.normal(0, spread, (H, W)).astype(np.float32)
    return mapx, mapy


@pytest.mark.parametrize('dtype', [np.uint8, np.uint16, np.float32, np.float64])
@pytest.mark.parametrize('border', [0, 7, 1234.5])
def test_remap_model_vs_cv2(dtype, border):
    rng = np.random.default_rng(3)
    H, W = 70, 90
    if np.dtype(dtype).kind == 'u':
        src = rng.integers(0, np.iinfo(dtype).max + 1, (H, W)).astype(dtype)
    else:
        src = ((rng.random((H, W)) - 0.3) * 4000).astype(dtype)
    mapx, mapy = _rand_maps(rng, H, W, 4.0)
    # far outside, exactly on the rim, ties at .5/32, NaN / inf coordinates
    mapx[0, :6] = [-1.0, -0.5, W - 1, W - 0.5, W, 1e9]
    mapy[1, :4] = [-1.0, H - 1, H, -1e9]
    mapx[2, :3] = [np.nan, np.inf, -np.inf]
    mapx[3, :4] = np.float32(10) + np.array([0.5, 1.5, 2.5, 3.5], np.float32) / np.float32(32)
    ref = cv2.remap(src, mapx, mapy, cv2.INTER_LINEAR, borderMode=cv2.BORDER_CONSTANT, borderValue=border)
    got = models.remap_model(src, mapx, mapy, border)
    assert got.dtype == ref.dtype
    assert np.array_equal(got, ref)


@pytest.mark.parametrize('shape,params', [((300, 400), None), ((501, 665), 'realistic'), ((256, 256), 'strong')])
def test_map_model_vs_cv2(shape, params):
    H, W = shape
    p = synth.lens_moderate(H, W) if params is None else \
        (synth.lens_realistic() if params == 'realistic' else synth.lens_strong(H, W))
    K, d = synth.camera_matrix(p), synth.dist_coeffs(p)
    mapx, mapy, P, roi = refpath.undistort_rectify_map(K, d, W, H)
    mx, my = models.undistort_map_model(K, d, P, W, H)
    for a, b in zip(models.fixed_point_coords(mx, my), models.fixed_point_coords(mapx, mapy)):
        assert np.array_equal(a, b)
    assert np.abs(mx.astype(np.float64) - mapx).max() < 1e-3


def test_pointwise_model_is_float32_of_reference():
    rng = np.random.default_rng(5)
    H, W = 50, 60
    raw = rng.integers(0, 65536, (H, W)).astype(np.uint16)
    dark = synth.dark_map(H, W)
    flat = synth.flat_map(H, W, p_zero=0.01)
    x = refpath.to_float_image(raw)
    refpath.correct_dark_current(x, dark)
    refpath.correct_vignetting(x, flat)
    assert np.array_equal(models.pointwise_model(raw, dark, flat, nan_to_num=False), x.astype(np.float32))


def test_full_size_chain_smoke():
    """one 1024x1024 float32 frame (BASELINE.json configs[0]) through both layers."""
    H = W = 1024
    raw = synth.scene(H, W, 0, np.float32)
    dark, flat = synth.dark_map(H, W), synth.flat_map(H, W)
    p = synth.lens_moderate(H, W)
    K, d = synth.camera_matrix(p), synth.dist_coeffs(p)
    ref = refpath.correct(raw, dark, flat, (K, d), 0.1)
    _, _, P, _ = refpath.undistort_rectify_map(K, d, W, H)
    got, mask = models.correct_chain_f32(raw, dark, flat, 0.1, 3, (K, d, P))
    assert 0.001 < mask.mean() < 0.05
    assert np.abs(got - ref).max() / 4095.0 < 1e-5


def test_band_chain_equals_full_chain():
    """oracle/bands.py (what bench.py and the full-size GPU tests check against) reproduces the full-frame oracle chain
    row for row, for 3x3 and 5x5, at frame edges and interior bands, through a strong lens"""
    import cv2
    from imgprocessor_b200 import synth
    from oracle import bands
    H, W = 300, 416
    raw = synth.scene(H, W, 3, np.uint16)
    dark, flat = synth.dark_map(H, W), synth.flat_map(H, W, p_zero=1e-3)
    for lens in (synth.lens_moderate(H, W), synth.lens_strong(H, W)):
        K, d = synth.camera_matrix(lens), synth.dist_coeffs(lens)
        P, _ = cv2.getOptimalNewCameraMatrix(K, d, (W, H), 1, (W, H))
        mapx, mapy = cv2.initUndistortRectifyMap(K, d, None, P, (W, H), cv2.CV_32FC1)
        for size in (3, 5):
            full, _ = models.correct_chain_f32(raw, dark, flat, 0.1, size, mapxy=(mapx, mapy))
            k1, _ = models.median_threshold_model(models.pointwise_model(raw, dark, flat, True), 0.1, size)
            for r0, r1 in bands.default_bands(H, 40) + [(7, 19), (290, 300)]:
                assert np.array_equal(bands.chain_band(raw, dark, flat, mapx, mapy, r0, r1, 0.1, size), full[r0:r1])
                assert np.array_equal(bands.k1_band(raw, dark, flat, r0, r1, 0.1, size), k1[r0:r1])
    full, _ = models.correct_chain_f32(raw, dark, flat, 0, 3, mapxy=(mapx, mapy))
    assert np.array_equal(bands.chain_band(raw, dark, flat, mapx, mapy, 100, 140, 0, 3), full[100:140], equal_nan=True)
