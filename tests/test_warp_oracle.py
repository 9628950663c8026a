"""SURVEY §8 row f3 on the CPU: the warpPerspective model (oracle/warp.py) against the reference's own
PerspectiveCorrection outputs (tests/golden/perspective.npz) and live OpenCV; the kernels' arithmetic
(csrc/imgcorr_warp.cuh compiled with g++) against the model; host logic of the PerspectiveCorrection mirror."""
import numpy as np
import pytest

from conftest import load_golden
import emul
from oracle import refpath
from oracle import warp as W

cv2 = pytest.importorskip('cv2')

QUADS = [np.array([[0.5, 0.1, -20], [0.05, 0.7, -30], [1e-4, -2e-4, 1.0]]),
         np.eye(3) + np.array([[0, 0, 5.3], [0, 0, -7.77], [0, 0, 0]]),
         np.array([[1.3, -0.2, 12.5], [0.1, 1.1, -3.0], [-3e-4, 1e-4, 1.0]])]


def _img(dt, shape=(90, 130), seed=0):
    rng = np.random.default_rng(seed)
    a = rng.random(shape)
    if dt == np.uint8:
        return (a * 255).astype(np.uint8)
    return (a * 65535).astype(np.uint16) if dt == np.uint16 else a.astype(dt)


def test_tables_and_invert_match_opencv():
    lz, cu, _ = emul.warp_tables()
    assert np.array_equal(lz, W.lanczos4_table())
    assert np.array_equal(cu, W.cubic_table())
    # OpenCV's own tables, read out as impulse responses
    src = np.zeros((32, 32), np.float32)
    src[16, 16] = 1
    for k in range(32):
        M = np.array([[1, 0, -k / 32], [0, 1, 0], [0, 0, 1.0]])
        r = cv2.warpPerspective(src, M, (32, 32), flags=cv2.INTER_LANCZOS4)
        assert np.array_equal(np.array([r[16, 19 - c] for c in range(8)]), lz[k])
        r = cv2.warpPerspective(src, M, (32, 32), flags=cv2.INTER_CUBIC)
        assert np.array_equal(np.array([r[16, 17 - c] for c in range(4)]), cu[k])
    rng = np.random.default_rng(3)
    for i in range(100):
        M = rng.normal(size=(3, 3)) * rng.choice([1, 10, 1e-3])
        assert np.array_equal(cv2.invert(M)[1], W.invert3x3_cv(M))
        assert np.array_equal(emul.warp_tables(M)[2], W.invert3x3_cv(M))
    assert np.array_equal(W.invert3x3_cv(np.zeros((3, 3))), np.zeros((3, 3)))


@pytest.mark.parametrize('dt', [np.float32, np.float64, np.uint16, np.uint8])
@pytest.mark.parametrize('interp', ['lanczos4', 'cubic'])
def test_model_and_emul_match_opencv(dt, interp):
    img = _img(dt)
    flag = cv2.INTER_LANCZOS4 if interp == 'lanczos4' else cv2.INTER_CUBIC
    for M in QUADS:
        for dsize in ((130, 90), (37, 150), (200, 11)):
            for inv in (False, True):
                for border in (0.0, 1234.5):
                    ref = cv2.warpPerspective(img, M, dsize, flags=flag | (cv2.WARP_INVERSE_MAP if inv else 0),
                                              borderValue=border)
                    assert np.array_equal(W.warp_perspective_model(img, M, dsize, interp, inv, border), ref)
                    assert np.array_equal(emul.warp(img, M, dsize, interp, inv, border), ref)


def test_degenerate_homographies():
    img = _img(np.float32, (40, 50))
    for M in (np.zeros((3, 3)),                                           # singular: cv::invert gives zeros
              np.array([[1, 0, 0], [0, 1, 0], [0.05, 0, -1.0]]),          # W crosses zero inside the image
              np.array([[1e6, 0, 0], [0, 1e6, 0], [0, 0, 1.0]]),
              np.array([[1e-7, 0, 3], [0, 1e-7, 4], [0, 0, 1.0]])):
        for inv in (False, True):
            ref = cv2.warpPerspective(img, M, (64, 48), flags=cv2.INTER_LANCZOS4 | (cv2.WARP_INVERSE_MAP if inv else 0))
            assert np.array_equal(W.warp_perspective_model(img, M, (64, 48), 'lanczos4', inv), ref, equal_nan=True)
            assert np.array_equal(emul.warp(img, M, (64, 48), 'lanczos4', inv), ref, equal_nan=True)


def test_special_values():
    img = _img(np.float32, (40, 50))
    img[10, 10] = np.inf
    img[20, 30] = np.nan
    img[5, 40] = -np.inf
    M = QUADS[2]
    ref = cv2.warpPerspective(img, M, (60, 50), flags=cv2.INTER_LANCZOS4)
    assert np.array_equal(emul.warp(img, M, (60, 50)), ref, equal_nan=True)
    assert np.array_equal(W.warp_perspective_model(img, M, (60, 50)), ref, equal_nan=True)


def test_golden_reference_outputs():
    g = load_golden('perspective')
    scene = g['scene']
    ns = tuple(int(v) for v in g['new_size'])
    Hq = g['quad_homography']
    for tag, img in (('f64', scene), ('f32', scene.astype(np.float32)), ('u16', (scene * 65535).astype(np.uint16))):
        assert np.array_equal(refpath.perspective_correct(img, Hq, ns), g['quad_' + tag])
        assert np.array_equal(W.warp_perspective_model(img, Hq, ns[::-1]), g['quad_' + tag])
        assert np.array_equal(emul.warp(img, Hq, ns[::-1]), g['quad_' + tag])
    assert np.array_equal(refpath.perspective_uncorrect(g['quad_f64'], Hq), g['uncorrect_f64'])
    assert np.array_equal(W.warp_perspective_model(g['quad_f64'], Hq, ns[::-1], 'cubic', True), g['uncorrect_f64'])
    Hm = g['Hm']
    assert np.array_equal(W.warp_perspective_model(scene, Hm, scene.shape[::-1]), g['homography_f64'])
    assert np.array_equal(W.warp_perspective_model(scene.astype(np.float32), Hm, scene.shape[::-1]), g['homography_f32'])
    big = (scene.shape[1] + 50, scene.shape[0] + 30)
    assert np.array_equal(W.warp_perspective_model(scene.astype(np.float32), Hm, big, border_value=0.25),
                          g['homography_border_f32'])
    assert np.array_equal(emul.warp(scene.astype(np.float32), Hm, big, border=0.25), g['homography_border_f32'])


def test_tilt_factor_division_model():
    img = _img(np.uint16)
    tf = 0.5 + np.random.default_rng(5).random(img.shape)
    M = QUADS[0]
    ref = refpath.perspective_correct(img, M, (70, 110), tilt_factor=tf)
    assert ref.dtype == np.float64
    assert np.array_equal(W.warp_perspective_model(img, M, (110, 70), divide_by=tf), ref)


def test_sort_corners_and_quad_homography():
    from scipy.spatial import ConvexHull
    from imgprocessor_b200.camera.PerspectiveCorrection import sortCorners, genericCameraMatrix

    def restated(corners):                      # utils/sortCorners.py:8-47 for a convex quad
        corners = np.asarray(corners)
        c2 = corners[ConvexHull(corners).vertices]
        d = c2 - c2.mean(axis=0)
        bl = np.abs(2.356194490192345 + np.arctan2(d[:, 1], d[:, 0])).argmin()
        return c2[list(range(bl, 4)) + list(range(0, bl))]

    rng = np.random.default_rng(1)
    n = 0
    for t in range(500):
        q = rng.random((4, 2)) * 100
        if len(ConvexHull(q).vertices) != 4:
            with pytest.raises(ValueError):
                sortCorners(q)
            continue
        n += 1
        assert np.array_equal(sortCorners(q), restated(q))
    assert n > 200
    g = load_golden('perspective')
    assert np.array_equal(sortCorners(g['quad']), g['quad_sorted'])
    K = genericCameraMatrix((120, 160))
    assert K.dtype == np.float32 and K[0, 2] == 80 and K[1, 2] == 60


def test_random_homographies_emul_vs_opencv():
    """seeded sweep: random quads -> homographies (zoom, rotation, keystone), random sizes / dtypes / borders"""
    rng = np.random.default_rng(2024)
    for t in range(40):
        H, Wd = int(rng.integers(9, 70)), int(rng.integers(9, 90))
        dt = [np.uint8, np.uint16, np.float32, np.float64][t % 4]
        img = _img(dt, (H, Wd), seed=t)
        dw, dh = int(rng.integers(1, 100)), int(rng.integers(1, 80))
        quad = np.float32([[0, 0], [Wd, 0], [Wd, H], [0, H]]) + rng.normal(0, 0.15 * min(H, Wd), (4, 2)).astype(np.float32)
        dst = np.float32([[0, 0], [dw, 0], [dw, dh], [0, dh]])
        try:
            M = cv2.getPerspectiveTransform(quad, dst)
        except cv2.error:
            continue
        interp, flag = (('lanczos4', cv2.INTER_LANCZOS4), ('cubic', cv2.INTER_CUBIC))[t % 2]
        inv = bool((t // 2) % 2)
        border = float(rng.integers(0, 300))
        ref = cv2.warpPerspective(img, M, (dw, dh), flags=flag | (cv2.WARP_INVERSE_MAP if inv else 0), borderValue=border)
        assert np.array_equal(emul.warp(img, M, (dw, dh), interp, inv, border), ref.reshape(dh, dw), equal_nan=True), t


def test_perspective_correction_host_logic():
    """the mirror's host side without a GPU: argument conventions and the loud refusals"""
    from imgprocessor_b200.camera.PerspectiveCorrection import PerspectiveCorrection, sortCorners
    pc = PerspectiveCorrection((120, 160), new_size=(96, 136), border=4)
    g = load_golden('perspective')
    pc.setReference(g['quad'])
    assert np.array_equal(pc.quad, g['quad_sorted'])
    assert np.array_equal(pc.homography, g['quad_homography'])          # cv2.getPerspectiveTransform on the sorted quad
    assert pc.obj_points.shape == (4, 3) and pc.obj_points.dtype == np.float32
    area = pc.areaRatio
    assert area > 0.9 and area < 1.5
    pts = pc.correctPoints(np.float32(g['quad_sorted']))
    assert np.allclose(pts[0], [[4, 4], [132, 4], [132, 92], [4, 92]], atol=1e-3)   # the quad lands on the bordered target
    with pytest.raises(NotImplementedError):
        pc.setReference(np.zeros((50, 60)))                               # reference IMAGE: PatternRecognition path
    with pytest.raises(ValueError):
        sortCorners([[0, 0], [10, 0], [3, 2], [0, 10]])                   # concave
    with pytest.raises(NotImplementedError):
        PerspectiveCorrection((120, 160)).setReference(g['quad'])         # new_size not given
    pc2 = PerspectiveCorrection((120, 160), new_size=(96, 136), do_correctIntensity=True)
    with pytest.raises(NotImplementedError):
        pc2.tiltFactor()
    pc2.setTiltFactor(np.ones((120, 160)))
    assert pc2.tiltFactor().dtype == np.float64
    with pytest.raises(NotImplementedError):
        PerspectiveCorrection._cv2_opts({'borderMode': 1})
    assert PerspectiveCorrection._cv2_opts({'borderValue': (3.0, 0, 0, 0)}) == 3.0
