"""SURVEY §8 row f4 — calibration-map producers: oracle/producers.py against outputs of the unmodified reference
(tests/golden/producers.npz), and the GPU mirrors (kernel K5, K4, K1) against both."""
import numpy as np
import pytest

from conftest import load_golden
from oracle import producers


def test_oracle_flat_field_from_close_distance_matches_reference():
    g = load_golden('producers')
    imgs, bgs = list(g['ff_imgs']), list(g['ff_bgs'])
    assert np.array_equal(producers.flat_field_from_close_distance(imgs, bgs), g['ff_bglist'])
    assert np.array_equal(producers.flat_field_from_close_distance(imgs, 41.5), g['ff_bgnum'])
    f32 = producers.flat_field_from_close_distance([i.astype(np.float32) for i in imgs], [b.astype(np.float32) for b in bgs])
    assert np.array_equal(f32, g['ff_f32'])


def test_oracle_linear_regression_is_least_squares():
    """the restated fancytools regression (unpinned) is ordinary least squares: equals np.polyfit per pixel where nothing
    is masked, ignores masked samples, and getLinearityFunction's clean-up follows DarkCurrentMap.py:72-78"""
    rng = np.random.default_rng(2)
    x = np.array([1.0, 2.0, 4.0, 8.0, 16.0])
    y = 100.0 + rng.random((1, 6, 7)) * 5 + x[:, None, None] * (rng.random((1, 6, 7)) * 3) + rng.normal(0, 0.1, (5, 6, 7))
    a, b, e = producers.lin_regress_masked(x, y, np.zeros(y.shape, bool))
    for i in range(6):
        for j in range(7):
            p = np.polyfit(x, y[:, i, j], 1)
            assert abs(a[i, j] - p[0]) < 1e-9 and abs(b[i, j] - p[1]) < 1e-8
            assert abs(e[i, j] - np.sqrt(np.mean((y[:, i, j] - np.polyval(p, x)) ** 2))) < 1e-9
    y2 = y.copy()
    y2[4, 2, 3] = 70000.0                                       # saturated sample: masked
    a2, b2, _ = producers.lin_regress_masked(x, y2, y2 > 65535)
    p = np.polyfit(x[:4], y[:4, 2, 3], 1)
    assert abs(a2[2, 3] - p[0]) < 1e-9 and abs(b2[2, 3] - p[1]) < 1e-8
    flat = np.broadcast_to(np.array([50.0, 50.0004, 50.0008, 50.0016, 50.0032])[:, None, None], (5, 2, 2)).copy()
    off, asc, _ = producers.get_linearity_function(x, flat, min_ascent=0.001)
    assert (asc == 0).all() and np.allclose(off, 50.0 + 0.5 * (1 + 16) * 0.0002, atol=1e-6)


gpu = pytest.mark.gpu


@gpu
def test_flat_field_from_close_distance_gpu():
    import torch
    if not torch.cuda.is_available():
        pytest.skip('needs a CUDA device')
    from imgprocessor_b200.camera.flatField import flatFieldFromCloseDistance
    g = load_golden('producers')
    imgs, bgs = list(g['ff_imgs']), list(g['ff_bgs'])
    out = flatFieldFromCloseDistance(imgs, bgs)
    assert out.dtype == np.float64 and np.array_equal(out, g['ff_bglist'])
    assert np.array_equal(flatFieldFromCloseDistance(imgs, 41.5), g['ff_bgnum'])
    assert np.array_equal(flatFieldFromCloseDistance([i.astype(np.float32) for i in imgs], [b.astype(np.float32) for b in bgs]), g['ff_f32'])
    with pytest.raises(ValueError):                              # gray frames: the reference's toGray raises as well
        flatFieldFromCloseDistance([i[..., 0] for i in imgs], 0.0)
    with pytest.raises(NotImplementedError):
        flatFieldFromCloseDistance(imgs)


@gpu
def test_dark_current_producers_gpu():
    import torch
    if not torch.cuda.is_available():
        pytest.skip('needs a CUDA device')
    from imgprocessor_b200.camera import DarkCurrentMap as dcm
    g = load_golden('producers')
    frames, times = g['dc_frames'], list(g['dc_times'])
    assert np.array_equal(dcm.averageSameExpTimes(list(frames[3:7])), g['dc_avg_t4'])       # K4, nStd = 3, estimated NLF
    xs, av = dcm.getDarkCurrentAverages(times, list(frames))
    assert xs == list(g['dc_x']) and av.dtype == g['dc_averages'].dtype and np.array_equal(av, g['dc_averages'])
    # the per-pixel line (K5) against the restated regression — bit for bit, incl. masked (saturated) samples, degenerate
    # pixels (all samples masked but one -> NaN ascent -> 0) and the min_ascent clean-up
    rng = np.random.default_rng(3)
    x = [0.5, 1.0, 2.0, 4.0, 8.0, 16.0]
    H, W = 70, 90
    stack = np.stack([np.clip(np.rint(100 + 10 * rng.random((H, W)) + t * 30 * rng.random((H, W)) + rng.normal(0, 2, (H, W))), 0, 65535)
                      for t in x]).astype(np.uint16)
    stack[3:, 5, 5] = 65535
    stack[1:, 6, 6] = 65535
    stack[:, 7, 7] = 200
    for dt, mx in ((np.uint16, 65534), (np.float64, 65534), (np.float32, 300.0)):
        s = stack.astype(dt)
        off, asc, err = dcm.getLinearityFunction(x, s, mxIntensity=mx, min_ascent=0.001)
        woff, wasc, werr = producers.get_linearity_function(x, s, mx, 0.001)
        assert np.array_equal(asc, wasc) and np.array_equal(off, woff, equal_nan=True) and np.array_equal(err, werr, equal_nan=True)
    off, asc, err = dcm.getDarkCurrentFunction(times, list(frames))
    woff, wasc, werr = producers.get_linearity_function(list(g['dc_x']), g['dc_averages'])
    assert np.array_equal(asc, wasc) and np.array_equal(off, woff) and np.array_equal(err, werr)
    assert 1.5 < np.median(asc) < 3.5                          # the synthetic series grows by 2.5 counts per unit time


@gpu
def test_lens_map_utilities_gpu():
    """LensDistortion.getDistortRectifyMap / getShift / getDeflection / distortImage (LensDistortion.py:332-340, 382-402)
    against the reference's outputs; they derive from the analytic float32 maps, which agree with OpenCV's to 1 ulp on a
    vanishing fraction of entries (DESIGN.md "map precision"), hence the tolerances"""
    import torch
    if not torch.cuda.is_available():
        pytest.skip('needs a CUDA device')
    from imgprocessor_b200.camera import LensDistortion
    g = load_golden('producers')
    H, W = g['lens_img8'].shape
    lens = LensDistortion({'cameraMatrix': g['lens_K'], 'distortionCoeffs': g['lens_dist'], 'shape': (H, W)})
    mx, my = lens.getDistortRectifyMap(W, H)
    assert mx.dtype == np.float32 and np.abs(mx - g['lens_dmapx']).max() < 1e-3 and np.abs(my - g['lens_dmapy']).max() < 1e-3
    assert (mx != g['lens_dmapx']).mean() < 5e-3
    assert np.abs(lens.getShift(W, H) - g['lens_shift']).max() < 1e-3
    ux, uy = lens.getDeflection(W, H)
    assert np.allclose(ux, g['lens_ux'], rtol=1e-3) and np.allclose(uy, g['lens_uy'], rtol=1e-3)
    d8 = lens.distortImage(g['lens_img8'])
    assert d8.dtype == np.uint8 and (d8 != g['lens_distort8']).mean() < 2e-3
    df = lens.distortImage(g['lens_imgf'])
    assert df.dtype == np.float32 and np.abs(df - g['lens_distortf']).max() / 4095.0 < 1e-3
