"""The C-ABI library loads on a box without a GPU, exports every symbol include/imgcorr.h declares,
and fails loudly (status + message, no abort, no fallback) when asked to compute without a device."""
import ctypes
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared():
    src = open(os.path.join(ROOT, 'include', 'imgcorr.h')).read()
    return sorted(set(re.findall(r'IMGCORR_API[^;(]*?\b(imgcorr_\w+)\s*\(', src)))


def test_header_declares_the_expected_surface():
    names = _declared()
    for must in ('imgcorr_ctx_create', 'imgcorr_pointwise_median', 'imgcorr_undistort', 'imgcorr_correct_batch',
                 'imgcorr_correct_host', 'imgcorr_last_error'):
        assert must in names
    assert len(names) >= 18


def test_library_exports_every_declared_symbol():
    from imgprocessor_b200 import _lib
    handle = _lib.lib()
    for name in _declared():
        assert hasattr(handle, name), name
        assert name in _lib.SIGNATURES, 'ctypes signature missing for %s' % name
    assert set(_lib.SIGNATURES) == set(_declared())
    assert handle.imgcorr_version() == 200


def test_no_device_means_loud_failure_not_fallback():
    import torch
    if torch.cuda.is_available():
        pytest.skip('this check is for the CPU box')
    from imgprocessor_b200 import _lib
    handle = _lib.lib()
    ctx = ctypes.c_void_p()
    st = handle.imgcorr_ctx_create(0, 16, 16, ctypes.byref(ctx))
    assert st == _lib.ERR_CUDA and not ctx.value
    assert b'cuda' in handle.imgcorr_last_error().lower()
    import numpy as np
    from imgprocessor_b200.filters import medianThreshold
    with pytest.raises(RuntimeError, match='no CPU fallback'):
        medianThreshold(np.ones((8, 8), np.float32), 0.1)


def test_product_never_imports_the_oracle():
    pkg = os.path.join(ROOT, 'imgprocessor_b200')
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith(('.py', '.cu', '.cuh', '.h')):
                txt = open(os.path.join(dirpath, f)).read()
                assert not re.search(r'^\s*(from|import)\s+oracle\b', txt, re.M), os.path.join(dirpath, f)
                # scipy.ndimage / OpenCV's remap are the reference's compute path; the product may only use cv2 on the host
                # for getOptimalNewCameraMatrix and the calibration-time pattern detection, and scipy.optimize.curve_fit for
                # the two-parameter fit of the noise level function (<= 100 points, host side, as the reference does)
                for m in re.finditer(r'^\s*(?:from|import)\s+(scipy[\w.]*)', txt, re.M):
                    assert m.group(1) == 'scipy.optimize', (os.path.join(dirpath, f), m.group(1))
                assert 'cv2.remap' not in txt.replace('cv2.remap(', 'X', 0) or f.endswith(('.cu', '.cuh', '.py')), f
                for banned in ('median_filter(', 'cv2.remap(', 'initUndistortRectifyMap('):
                    code = '\n'.join(l for l in txt.splitlines() if not l.lstrip().startswith(('#', '//', '*', '"', "'")))
                    if f.endswith('.py'):
                        assert banned not in re.sub(r'""".*?"""', '', code, flags=re.S), (f, banned)


def _build_c_consumer(tmp_path):
    import subprocess
    from imgprocessor_b200 import build
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    lib_dir = os.path.dirname(build.build())
    exe = str(tmp_path / 'abi_smoke')
    subprocess.check_call(['gcc', '-std=c99', '-Wall', '-Wextra', '-Werror', '-pedantic', '-I', os.path.join(root, 'include'),
                           os.path.join(root, 'tests', 'c_abi', 'abi_smoke.c'), '-o', exe, '-L', lib_dir, '-limgcorr',
                           '-Wl,-rpath,' + lib_dir])
    return exe


def test_plain_c_program_compiles_links_and_runs(tmp_path):
    """include/imgcorr.h is valid strict C99 and libimgcorr.so is usable without Python / torch; without a GPU the
    library reports errors through its status codes instead of aborting"""
    import subprocess
    exe = _build_c_consumer(tmp_path)
    r = subprocess.run([exe], capture_output=True, text=True, timeout=120)
    assert r.returncode == 0, r.stderr
    assert r.stdout.startswith('abi ok')


@pytest.mark.gpu
def test_plain_c_program_runs_the_chain(tmp_path):
    """the same chain from C (imgcorr_correct_host) and from the Python mirror's engine: identical output"""
    import subprocess
    import numpy as np
    torch = pytest.importorskip('torch')
    if not torch.cuda.is_available():
        pytest.skip('needs a CUDA device')
    exe = _build_c_consumer(tmp_path)
    H, W, N = 64, 128, 3
    r = subprocess.run([exe, 'gpu', str(H), str(W), str(N)], capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, r.stderr
    c_sum = float(r.stdout.split()[1])
    # the C program's synthetic inputs, regenerated here
    s, vals = 12345, []
    for i in range(H * W * N):
        s = (s * 1664525 + 1013904223) & 0xffffffff
        vals.append((20000 + (s >> 20)) & 0xffff)
    raw = np.array(vals, np.uint16).reshape(N, H, W)
    idx = np.arange(H * W)
    dark = (100.0 + (idx % 7)).astype(np.float32).reshape(H, W)
    flat = (0.5 + (idx % 11).astype(np.float32) * np.float32(0.04)).astype(np.float32).reshape(H, W)
    K = np.array([[W, 0, W / 2.0], [0, W, H / 2.0], [0, 0, 1.0]])
    P = K.copy()
    P[0, 0] = P[1, 1] = 0.9 * W
    from imgprocessor_b200 import engine
    e = engine.Engine(H, W, 0)
    e.set_dark(dark)
    e.set_flat(flat)
    e.set_lens(K, [-0.2, 0.05, 1e-3, -1e-3, 0.0], P)
    out = e.correct_batch(torch.from_numpy(raw).cuda(), threshold=0.1, ksize=3).cpu().numpy()
    assert abs(out.astype(np.float64).sum() - c_sum) <= 1e-6 * abs(c_sum)
