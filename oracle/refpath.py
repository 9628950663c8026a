"""Restatement of the reference's per-frame chain, float64 end to end, calling
the same third-party natives the reference calls.

TEST INFRASTRUCTURE ONLY (see oracle/__init__.py).  This is also what the
``cpu_baseline`` and ``--impl reference`` legs of bench.py time: /root/reference
is a Python package that cannot travel to the GPU box, and every hot loop of its
path lives in scipy / OpenCV / numpy, so this restatement runs the same native
code in the same order on the same dtypes.

Every function cites the reference lines it follows (radjkarl/imgProcessor
0.2.5, paths relative to /root/reference/imgProcessor/).
"""
import numpy as np


def to_float_image(image):
    """camera/CameraCalibration.py:408-410 -> imgIO.py:39-73,17-22:
    ``np.asfarray(imread(x, dtype=np.float))`` == a float64 COPY; the caller's
    array is never mutated."""
    return np.array(image, dtype=np.float64, copy=True)


def calc_dark_current(entry, exposure_time, depth_bits=16):
    """camera/CameraCalibration.py:504-518.  ``entry`` is one item of
    coeffs['dark current']: a *tuple* entry evaluates offs + ascent*t and clips
    to 2**depth-1; any other entry (the list that addDarkCurrent builds) returns
    its data field untouched."""
    if type(entry) == tuple:
        offs, ascent = entry[2]
        bg = offs + ascent * exposure_time
        mx = 2 ** depth_bits - 1
        with np.errstate(invalid='ignore'):
            bg[bg > mx] = mx
        return bg
    return entry[2]


def correct_dark_current(image, bg):
    """camera/CameraCalibration.py:502 — in-place float64 subtract."""
    image -= bg
    return image


def correct_vignetting(image, flat):
    """camera/CameraCalibration.py:520-526 — divide only where flat != 0."""
    i = flat != 0
    image[i] /= flat[i]
    return image


def median_threshold(img, threshold=0.1, size=3, condition='>', copy=True):
    """filters/medianThreshold.py:7-30."""
    from scipy.ndimage import median_filter
    indices = None
    if threshold > 0:
        blur = np.asarray(median_filter(img, size=size), dtype=np.float64)
        with np.errstate(divide='ignore', invalid='ignore', over='ignore'):
            rel = abs((img - blur) / blur)
            indices = rel > threshold if condition == '>' else rel < threshold
        if copy:
            img = img.copy()
        img[indices] = blur[indices]
    return img, indices


def correct_artefacts(image, threshold):
    """camera/CameraCalibration.py:556-563 — nan_to_num (new array) then the
    in-place 3x3 medianThreshold."""
    image = np.nan_to_num(image)
    median_threshold(image, threshold, copy=False)
    return image


def undistort_rectify_map(camera_matrix, dist_coeffs, width, height):
    """camera/LensDistortion.py:342-358.  Returns (mapx, mapy, P, roi)."""
    import cv2
    cam = np.asarray(camera_matrix, np.float64)
    d = np.asarray(dist_coeffs, np.float64)
    P, roi = cv2.getOptimalNewCameraMatrix(cam, d, (width, height), 1, (width, height))
    mapx, mapy = cv2.initUndistortRectifyMap(cam, d, None, P, (width, height), cv2.CV_32FC1)
    return mapx, mapy, P, roi


def lens_correct(image, camera_matrix, dist_coeffs, keep_size=False, border_value=0):
    """camera/LensDistortion.py:316-330 — dtype of ``image`` is preserved by cv2.remap."""
    import cv2
    h, w = image.shape[:2]
    mapx, mapy, _, roi = undistort_rectify_map(camera_matrix, dist_coeffs, w, h)
    out = cv2.remap(image, mapx, mapy, cv2.INTER_LINEAR,
                    borderMode=cv2.BORDER_CONSTANT, borderValue=border_value)
    if not keep_size:
        xx, yy, ww, hh = roi
        out = out[yy:yy + hh, xx:xx + ww]
    return out


def correct(image, dark=None, flat=None, lens=None, threshold=0.1, keep_size=True):
    """camera/CameraCalibration.py:351-459, single-frame branch, with the
    calibration lookup already resolved to arrays: ``dark`` the background image
    (:500-502), ``flat`` the flat-field array (:521-526), ``lens`` a
    (cameraMatrix, distortionCoeffs) pair (:565-575).  Stage errors are NOT
    swallowed here (the reference prints and continues, :418-452); the host
    mirror in imgprocessor_b200 reproduces that policy, the oracle wants to see
    them."""
    image = to_float_image(image)
    if dark is not None:
        correct_dark_current(image, dark)
    if flat is not None:
        correct_vignetting(image, flat)
    if threshold > 0:
        image = correct_artefacts(image, threshold)
    if lens is not None:
        image = lens_correct(image, lens[0], lens[1], keep_size=keep_size)
    return image


# ---- SURVEY §8 row f3 -------------------------------------------------------------------
def perspective_correct(image, homography, new_size, tilt_factor=None, **cv2_opts):
    """PerspectiveCorrection.correct (camera/PerspectiveCorrection.py:380-406) for a fixed homography:
    optional ``np.asfarray(img) / tiltFactor`` (:394-400), then
    ``cv2.warpPerspective(img, h, newBorders[::-1], flags=cv2.INTER_LANCZOS4, **cv2_opts)`` (:401-405).
    new_size = (sizey, sizex) as in the reference's constructor."""
    import cv2
    img = np.asarray(image)
    if tilt_factor is not None:
        img = np.asarray(img, dtype=np.float64) / tilt_factor
    return cv2.warpPerspective(img, np.asarray(homography, np.float64), tuple(new_size)[::-1], flags=cv2.INTER_LANCZOS4,
                               **cv2_opts)


def perspective_uncorrect(image, homography):
    """PerspectiveCorrection.uncorrect (camera/PerspectiveCorrection.py:374-378)."""
    import cv2
    img = np.asarray(image)
    return cv2.warpPerspective(img, np.asarray(homography, np.float64), img.shape[:2][::-1],
                               flags=cv2.INTER_CUBIC | cv2.WARP_INVERSE_MAP)
