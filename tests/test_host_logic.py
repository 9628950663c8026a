"""Host-side mirror of the reference's calibration store (no GPU needed): date ordering and lookup,
light-spectrum fallback, pickle round trip, shape checks — against the behaviour pinned from the
unmodified reference (tests/golden/correct_u16_dates.npz)."""
import numpy as np
import pytest

from imgprocessor_b200.camera import CameraCalibration, LensDistortion
from imgprocessor_b200.camera.CameraCalibration import _getFromDate


def test_date_ordering_and_lookup():
    cal = CameraCalibration()
    a, b, c = (np.full((4, 4), v, np.float32) for v in (1, 2, 3))
    cal.addDarkCurrent(a, date='01 Jan 15 - 10:00')
    cal.addDarkCurrent(c, date='01 Jan 17 - 10:00')
    cal.addDarkCurrent(b, date='01 Jan 16 - 10:00')
    assert [e[2][0, 0] for e in cal.coeffs['dark current']] == [3, 2, 1]          # newest first
    pick = lambda d: cal.calcDarkCurrent(None, d)[0, 0]
    assert pick(None) == 3 and pick('garbage') == 3 and pick('01 Jan 18 - 00:00') == 3
    assert pick('01 Jun 16 - 00:00') == 3          # the reference returns the next-newer calibration
    assert pick('01 Jun 15 - 00:00') == 2
    assert pick('01 Jan 14 - 00:00') == 1
    assert cal.dates('dark current') == ['01 Jan 17 - 10:00', '01 Jan 16 - 10:00', '01 Jan 15 - 10:00']
    with pytest.raises(IndexError):
        _getFromDate([], None)


def test_shape_check_and_light_fallback(capsys):
    cal = CameraCalibration()
    cal.addFlatField(np.ones((4, 6), np.float32), light_spectrum='IR')
    assert cal.coeffs['shape'] == (4, 6) and cal.coeffs['light spectra'] == ['IR']
    with pytest.raises(Exception, match='array shapes are different'):
        cal.addDarkCurrent(np.zeros((6, 4), np.float32))
    got = cal.getCoeff('flat field', 'visible')
    assert got[2].shape == (4, 6)
    assert 'no calibration found for [visible] - using [IR] instead' in capsys.readouterr().out
    assert cal.getCoeff('lens', 'visible') is None
    cal.transpose()
    assert cal.coeffs['shape'] == (6, 4) and cal.coeffs['flat field']['IR'][0][2].shape == (6, 4)


def test_pickle_round_trip_and_lens_store(tmp_path):
    cal = CameraCalibration()
    cal.setCamera('cam', 12)
    cal.addDarkCurrent(np.arange(12, dtype=np.float32).reshape(3, 4))
    l = LensDistortion()
    l.setCameraParams(100, 101, 2, 1.5, -0.1, 0.01, 0.001, 1e-3, -1e-3)
    assert l.getCameraParams() == (100, 101, 2, 1.5, -0.1, 0.01, 0.001, 1e-3, -1e-3)
    assert LensDistortion()._coeffs is not LensDistortion()._coeffs          # no shared mutable default
    cal.addLens(l)
    p = cal._correctPath(str(tmp_path / 'x'))
    assert p.endswith('.cal')
    cal.saveToFile(str(tmp_path / 'x'))
    back = CameraCalibration.loadFromFile(str(tmp_path / 'x'))
    assert back.coeffs['name'] == 'cam' and back.coeffs['depth'] == 12
    assert np.array_equal(back.coeffs['dark current'][0][2], cal.coeffs['dark current'][0][2])
    stored = back.coeffs['lens']['visible'][0][2]
    assert np.array_equal(stored['distortionCoeffs'], [[-0.1, 0.01, 1e-3, -1e-3, 0.001]])
    fn = l.writeToFile(str(tmp_path / 'lens'))
    l2 = LensDistortion()
    l2.readFromFile(fn)
    assert np.array_equal(l2.coeffs['cameraMatrix'], l.coeffs['cameraMatrix'])


def test_lens_calibration_host_side():
    cv2 = pytest.importorskip('cv2')
    # synthetic chessboard views through a known camera: calibrateCamera must recover it
    rng = np.random.default_rng(0)
    K = np.array([[800., 0, 320], [0, 800., 240], [0, 0, 1]])
    dist = np.array([-0.2, 0.05, 1e-3, -1e-3, 0.0])
    l = LensDistortion()
    l.calibrate(board_size=(7, 5), method='Manual')
    l.setImgShape((480, 640))
    obj = l._mkObjPoints((7, 5)) * 30.0
    l.objp = obj
    for _ in range(12):
        rvec = rng.normal(0, 0.25, 3)
        tvec = np.array([-90 + rng.normal(0, 20), -60 + rng.normal(0, 20), 500 + rng.normal(0, 50)])
        pts, _ = cv2.projectPoints(obj, rvec, tvec, K, dist)
        l.addPoints(pts.reshape(-1, 2))
    c = l.coeffs
    assert c['reprojectionError'] < 1e-2
    assert np.allclose(c['cameraMatrix'], K, rtol=2e-3, atol=1.0)
    assert np.allclose(np.ravel(c['distortionCoeffs'])[:2], dist[:2], atol=2e-2)


def test_parse_cpulist_and_numa_binding(tmp_path, monkeypatch):
    from imgprocessor_b200 import sharding
    assert sharding.parse_cpulist('0-3,8,10-11\n') == [0, 1, 2, 3, 8, 10, 11]
    assert sharding.parse_cpulist('') == []
    # without a CUDA device the binding reports the failure and leaves the process alone
    import os
    before = os.sched_getaffinity(0)
    info = sharding.bind_host_to_gpu(0, sysfs=str(tmp_path))
    assert info['bound'] is False and os.sched_getaffinity(0) == before


def test_calibration_upload_token_sees_in_place_changes():
    from imgprocessor_b200.camera.CameraCalibration import _map_token
    a = np.random.default_rng(0).random((300, 400)).astype(np.float32)
    t0 = _map_token('dark', a)
    assert t0 == _map_token('dark', a) and t0 != _map_token('flat', a)
    a *= 2                                       # in place: same object, new content
    assert _map_token('dark', a) != t0
    big = np.zeros((1500, 2048), np.float32)     # every byte counts: one patched hot pixel in a large map, anywhere
    t1 = _map_token('dark', big)
    for pos in ((0, 0), (1499, 2047), (777, 1023), (3, 5)):
        big[pos] = 1e-3
        t2 = _map_token('dark', big)
        assert t2 != t1, pos
        t1 = t2
    assert _map_token('dark', big.copy()) == t1  # a different object with the same content needs no new upload
    ro = big.copy()
    ro.setflags(write=False)                     # read-only owner: identity-keyed fast path, stable across calls
    assert _map_token('dark', ro) == _map_token('dark', ro) and _map_token('dark', ro)[1] == 'ro'
    view = big[:]                                # a view of a writeable array is NOT safe: full fingerprint
    view.setflags(write=False)
    assert _map_token('dark', view)[1] != 'ro'
    assert _map_token('dark', 3.0) is None
    b = a[::2]                                   # non-contiguous views work too
    assert _map_token('dark', b) == _map_token('dark', b)
