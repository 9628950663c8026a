"""SURVEY §8 rows a11 / f1 on the CPU: the STE restatement (oracle/ste.py) against outputs of the reference's own
SingleTimeEffectDetection / removeSinglePixels / boundedFunction (tests/golden/ste.npz; MaskedMovingAverage restated,
see the oracle's header), and the kernels' arithmetic (csrc/imgcorr_ste.cuh compiled with g++) against the oracle."""
import numpy as np
import pytest

from conftest import load_golden
import emul
from oracle import ste


def test_oracle_matches_reference_outputs():
    g = load_golden('ste')
    fr, nlf = g['frames'], tuple(g['nlf'])
    for n in (2, 3, 5):
        a, m = ste.ste_average(list(fr[:n]), nlf, 4, True)
        assert np.array_equal(a, g['noSTE_%d' % n]) and np.array_equal(m, g['mask_%d' % n])
        assert m.sum() > 20                                      # the fixture does contain STEs
    assert np.array_equal(ste.ste_average(list(fr.astype(np.float32)), nlf, 3), g['noSTE_f32_nstd3'])
    assert np.array_equal(ste.remove_single_pixels(g['rsp_in']), g['rsp_out'])
    assert g['rsp_out'].sum() < g['rsp_in'].sum()
    assert np.array_equal(ste.bounded_function(g['bf_x'], *nlf), g['bf_y'], equal_nan=True)


def test_emul_matches_golden_and_oracle():
    g = load_golden('ste')
    fr, nlf = g['frames'], tuple(g['nlf'])
    for n in (2, 3, 5):
        a, m = emul.ste(fr[:n], nlf, 4, True)
        assert np.array_equal(a, g['noSTE_%d' % n]) and np.array_equal(m, g['mask_%d' % n])
    assert np.array_equal(emul.ste(fr.astype(np.float32), nlf, 3), g['noSTE_f32_nstd3'])
    rng = np.random.default_rng(9)
    for shape in ((1, 1), (1, 7), (5, 1), (9, 33), (40, 70)):
        for n in (2, 4):
            f = rng.normal(500, 30, (n,) + shape)
            f[rng.random(f.shape) < 0.05] += 400
            for coeff in ((5.0, 10.0, 1.2), (0.0, 1e9, 1.0), (3.0, -50.0, 0.0)):   # sqrt of negatives -> minY floor
                a, m = emul.ste(f, coeff, 3.5, True)
                ra, rm = ste.ste_average(list(f), coeff, 3.5, True)
                assert np.array_equal(a, ra) and np.array_equal(m, rm)


def test_remove_single_pixels_order_independence():
    # the reference's in-place raster scan == the vectorised form (checked against the numba original in the golden
    # file; here: a brute-force sequential restatement on random masks, including borders)
    rng = np.random.default_rng(2)
    for t in range(20):
        m = rng.random((12, 17)) > 0.8
        seq = m.copy()
        for i in range(12):
            for j in range(17):
                if seq[i, j]:
                    nb = seq[max(0, i - 1):i + 2, max(0, j - 1):j + 2].sum() - 1
                    if nb == 0:
                        seq[i, j] = False
        assert np.array_equal(seq, ste.remove_single_pixels(m))
