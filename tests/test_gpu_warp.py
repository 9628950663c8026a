"""SURVEY §8 row f3 on the GPU: kernel K3 (csrc/k3_warp.cu) through the C ABI against OpenCV-exact arithmetic
(oracle/warp.py, oracle/refpath.py) and the golden outputs of the reference's PerspectiveCorrection.
Bar: bit-exact for uint16 / float32 / float64, Lanczos4 and bicubic, any border value."""
import contextlib
import io

import numpy as np
import pytest

from conftest import load_golden
from oracle import refpath
from oracle import warp as W

pytestmark = pytest.mark.gpu

torch = pytest.importorskip('torch')
cv2 = pytest.importorskip('cv2')


@pytest.fixture(scope='module', params=[0, 1, 2], ids=['auto', 'l1', 'tiles'])
def eng(request):
    if not torch.cuda.is_available():
        pytest.skip('needs a CUDA device')
    from imgprocessor_b200 import engine, _lib
    e = engine.get_engine(8, 8)             # K3 takes free frame shapes; the context supplies the device
    # 2 = shared-memory tiles forced: the fixture falls back per call where a shape is not eligible
    e.set_option(_lib.OPT_K3_VARIANT, request.param)
    yield _Eng(e, request.param)
    e.set_option(_lib.OPT_K3_VARIANT, 0)


MATS = [np.array([[0.5, 0.1, -20], [0.05, 0.7, -30], [1e-4, -2e-4, 1.0]]),
        np.eye(3) + np.array([[0, 0, 5.3], [0, 0, -7.77], [0, 0, 0]]),
        np.array([[1.3, -0.2, 12.5], [0.1, 1.1, -3.0], [-3e-4, 1e-4, 1.0]])]


class _Eng(object):
    """engine wrapper: with the tiles forced (param 2) a call that is not eligible for them (float64, rows that are
    not 16-byte multiples) is repeated on the automatic path"""

    def __init__(self, e, param):
        self.e, self.param = e, param

    def set_option(self, k, v):
        self.e.set_option(k, v)

    def warp_perspective(self, *a, **k):
        from imgprocessor_b200 import _lib
        try:
            return self.e.warp_perspective(*a, **k)
        except _lib.ImgcorrError as err:
            if self.param != 2 or 'not eligible' not in str(err):
                raise
            self.e.set_option(_lib.OPT_K3_VARIANT, 0)
            try:
                return self.e.warp_perspective(*a, **k)
            finally:
                self.e.set_option(_lib.OPT_K3_VARIANT, 2)


def _img(dt, shape, seed=0):
    a = np.random.default_rng(seed).random(shape)
    if dt == np.uint8:
        return (a * 255).astype(np.uint8)
    return (a * 65535).astype(np.uint16) if dt == np.uint16 else a.astype(dt)


def _dev(a):
    return torch.from_numpy(np.ascontiguousarray(a)).cuda()


@pytest.mark.parametrize('dt', [np.float32, np.float64, np.uint16, np.uint8])
@pytest.mark.parametrize('interp', ['lanczos4', 'cubic'])
def test_k3_bit_exact(eng, dt, interp):
    img = _img(dt, (90, 130))
    flag = cv2.INTER_LANCZOS4 if interp == 'lanczos4' else cv2.INTER_CUBIC
    for M in MATS:
        for dsize in ((130, 90), (37, 150), (200, 11), (1, 1)):
            for inv in (False, True):
                for border in (0.0, 1234.5):
                    ref = cv2.warpPerspective(img, M, dsize, flags=flag | (cv2.WARP_INVERSE_MAP if inv else 0),
                                              borderValue=border)
                    got = eng.warp_perspective(_dev(img), M, dsize, interp, inv, border).cpu().numpy()
                    assert np.array_equal(got, ref.reshape(got.shape)), (dt, interp, dsize, inv, border)
                    assert np.array_equal(got, W.warp_perspective_model(img, M, dsize, interp, inv, border))


def test_k3_tiny_sources(eng):
    # sources smaller than the tap window: every pixel takes the border arithmetic
    for shape in ((1, 1), (3, 5), (7, 7), (8, 8), (2, 40)):
        img = _img(np.float32, shape, 2)
        M = np.array([[1.1, 0.05, 0.4], [-0.02, 0.9, 0.3], [0, 0, 1.0]])
        for interp, flag in (('lanczos4', cv2.INTER_LANCZOS4), ('cubic', cv2.INTER_CUBIC)):
            ref = cv2.warpPerspective(img, M, (23, 17), flags=flag, borderValue=0.5)
            got = eng.warp_perspective(_dev(img), M, (23, 17), interp, False, 0.5).cpu().numpy()
            assert np.array_equal(got, ref)


def test_k3_degenerate_and_special_values(eng):
    img = _img(np.float32, (40, 50))
    for M in (np.zeros((3, 3)), np.array([[1, 0, 0], [0, 1, 0], [0.05, 0, -1.0]]),
              np.array([[1e6, 0, 0], [0, 1e6, 0], [0, 0, 1.0]]), np.array([[1e-7, 0, 3], [0, 1e-7, 4], [0, 0, 1.0]])):
        for inv in (False, True):
            ref = cv2.warpPerspective(img, M, (64, 48), flags=cv2.INTER_LANCZOS4 | (cv2.WARP_INVERSE_MAP if inv else 0))
            got = eng.warp_perspective(_dev(img), M, (64, 48), 'lanczos4', inv).cpu().numpy()
            assert np.array_equal(got, ref, equal_nan=True)
    img[10, 10] = np.inf
    img[20, 30] = np.nan
    img[5, 40] = -np.inf
    ref = cv2.warpPerspective(img, MATS[2], (60, 50), flags=cv2.INTER_LANCZOS4)
    got = eng.warp_perspective(_dev(img), MATS[2], (60, 50)).cpu().numpy()
    assert np.array_equal(got, ref, equal_nan=True)


def test_k3_batch_and_division(eng):
    M = MATS[0]
    for dt in (np.uint16, np.uint8):
        frames = np.stack([_img(dt, (120, 160), s) for s in range(5)])
        got = eng.warp_perspective(_dev(frames), M, (140, 100)).cpu().numpy()
        for i in range(5):
            assert np.array_equal(got[i], cv2.warpPerspective(frames[i], M, (140, 100), flags=cv2.INTER_LANCZOS4))
    frames = np.stack([_img(np.uint16, (120, 160), s) for s in range(5)])
    tf = 0.5 + np.random.default_rng(5).random((120, 160))
    got = eng.warp_perspective(_dev(frames), M, (140, 100), divide_by=torch.from_numpy(tf)).cpu().numpy()
    assert got.dtype == np.float64
    for i in range(5):
        assert np.array_equal(got[i], refpath.perspective_correct(frames[i], M, (100, 140), tilt_factor=tf))


def test_k3_full_frame(eng):
    # BASELINE frame size, a mild keystone: against OpenCV itself (the model needs minutes at 12 Mpx)
    H, Wd = 3000, 4096
    img = _img(np.float32, (H, Wd), 7)
    quad = np.float32([[60, 40], [4040, 75], [4000, 2950], [30, 2900]])
    dst = np.float32([[0, 0], [Wd, 0], [Wd, H], [0, H]])
    M = cv2.getPerspectiveTransform(quad, dst)
    ref = cv2.warpPerspective(img, M, (Wd, H), flags=cv2.INTER_LANCZOS4)
    got = eng.warp_perspective(_dev(img), M, (Wd, H)).cpu().numpy()
    assert np.array_equal(got, ref)
    u16 = (img * 65535).astype(np.uint16)
    assert np.array_equal(eng.warp_perspective(_dev(u16), M, (Wd, H)).cpu().numpy(),
                          cv2.warpPerspective(u16, M, (Wd, H), flags=cv2.INTER_LANCZOS4))


def test_k3_rejects(eng):
    from imgprocessor_b200 import _lib
    with pytest.raises(TypeError):
        eng.warp_perspective(_dev(np.zeros((8, 8), np.int32)), np.eye(3), (8, 8))
    param, eng = eng.param, eng.e
    eng.set_option(_lib.OPT_K3_VARIANT, 2)
    try:
        with pytest.raises(_lib.ImgcorrError):      # rows of 130 float32 are not 16-byte multiples: tiles not eligible
            eng.warp_perspective(_dev(np.zeros((8, 130), np.float32)), np.eye(3), (8, 8))
        with pytest.raises(_lib.ImgcorrError):
            eng.warp_perspective(_dev(np.zeros((8, 128), np.float64)), np.eye(3), (8, 8))
        with pytest.raises(_lib.ImgcorrError):
            eng.warp_perspective(_dev(np.zeros((8, 128), np.uint8)), np.eye(3), (8, 8))
        out = eng.warp_perspective(_dev(_img(np.float32, (64, 128))), MATS[1], (128, 64)).cpu().numpy()
        assert np.array_equal(out, cv2.warpPerspective(_img(np.float32, (64, 128)), MATS[1], (128, 64), flags=cv2.INTER_LANCZOS4))
    finally:
        eng.set_option(_lib.OPT_K3_VARIANT, param)


@pytest.mark.parametrize('dt', [np.float32, np.uint16])
@pytest.mark.parametrize('interp', ['lanczos4', 'cubic'])
def test_k3_tiled_shapes(eng, dt, interp):
    # widths that are 16-byte multiples (tile path), homographies from identity to strong zoom-out / rotation
    # (tiles whose window does not fit the staged box fall back to global gathers)
    flag = cv2.INTER_LANCZOS4 if interp == 'lanczos4' else cv2.INTER_CUBIC
    th = 0.6
    mats = MATS + [np.eye(3), np.array([[np.cos(th), -np.sin(th), 80], [np.sin(th), np.cos(th), -40], [0, 0, 1.0]]),
                   np.array([[0.3, 0, 0], [0, 0.3, 0], [0, 0, 1.0]]), np.array([[3.0, 0, -100], [0, 3.0, -50], [0, 0, 1.0]])]
    for shape in ((96, 128), (200, 328), (64, 1024)):
        img = _img(dt, shape, 4)
        for M in mats:
            for dsize in ((shape[1], shape[0]), (77, 33)):
                ref = cv2.warpPerspective(img, M, dsize, flags=flag, borderValue=3.0)
                got = eng.warp_perspective(_dev(img), M, dsize, interp, False, 3.0).cpu().numpy()
                assert np.array_equal(got, ref), (shape, dsize, M)
    frames = np.stack([_img(dt, (96, 256), s) for s in range(11)])       # more frames than pipeline stages
    got = eng.warp_perspective(_dev(frames), MATS[0], (256, 96), interp).cpu().numpy()
    for i in range(11):
        assert np.array_equal(got[i], cv2.warpPerspective(frames[i], MATS[0], (256, 96), flags=flag))


def test_perspective_correction_api_golden():
    """the Python mirror against the unmodified reference's outputs"""
    if not torch.cuda.is_available():
        pytest.skip('needs a CUDA device')
    from imgprocessor_b200.camera import PerspectiveCorrection
    g = load_golden('perspective')
    scene = g['scene']
    ns = tuple(int(v) for v in g['new_size'])
    for tag, img in (('f64', scene), ('f32', scene.astype(np.float32)), ('u16', (scene * 65535).astype(np.uint16))):
        pc = PerspectiveCorrection(img.shape, new_size=ns, border=int(g['border']))
        pc.setReference(g['quad'])
        buf = io.StringIO()
        with contextlib.redirect_stdout(buf):
            out = pc.correct(img)
        assert buf.getvalue() == str(g['log'])
        assert out.dtype == img.dtype and np.array_equal(out, g['quad_' + tag])
        if tag == 'f64':
            assert np.array_equal(pc.quad, g['quad_sorted'])
            assert np.array_equal(pc.homography, g['quad_homography'])
            assert np.array_equal(pc.uncorrect(out), g['uncorrect_f64'])
    pc = PerspectiveCorrection(scene.shape, new_size=scene.shape)
    pc.setReference(g['Hm'])
    with contextlib.redirect_stdout(io.StringIO()):
        assert np.array_equal(pc.correct(scene), g['homography_f64'])
        assert np.array_equal(pc.correct(scene.astype(np.float32)), g['homography_f32'])
        t = pc.correct(torch.from_numpy(scene.astype(np.float32)).cuda())          # device tensors stay on the device
        assert t.is_cuda and np.array_equal(t.cpu().numpy(), g['homography_f32'])
    pc = PerspectiveCorrection(scene.shape, new_size=(scene.shape[0] + 30, scene.shape[1] + 50), cv2_opts={'borderValue': 0.25})
    pc.setReference(g['Hm'])
    with contextlib.redirect_stdout(io.StringIO()):
        assert np.array_equal(pc.correct(scene.astype(np.float32)), g['homography_border_f32'])
    # tilt-factor division + warp
    tf = 0.5 + np.random.default_rng(5).random(scene.shape)
    pc = PerspectiveCorrection(scene.shape, new_size=ns, do_correctIntensity=True)
    pc.setReference(g['Hm'])
    with pytest.raises(NotImplementedError):
        pc.correct(scene)
    pc.setTiltFactor(tf)
    with contextlib.redirect_stdout(io.StringIO()):
        out = pc.correct((scene * 65535).astype(np.uint16))
    assert np.array_equal(out, refpath.perspective_correct((scene * 65535).astype(np.uint16), g['Hm'], ns, tilt_factor=tf))
