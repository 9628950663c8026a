"""Ingest formats: the package's mirrors of imgProcessor.reader.RAW / elbin against what the reference's
own readers returned for the two committed files (tests/golden/readers.npz, made by make_golden.py), and the layout
helper the GPU path consumes."""
import os

import numpy as np
import pytest

from conftest import GOLDEN, load_golden
from imgprocessor_b200 import reader

RAW_FILE = os.path.join(GOLDEN, 'raw_be_u16_24x32.raw')
ELBIN_FILE = os.path.join(GOLDEN, 'stack_3x24x40.elbin')


def test_raw_reader_matches_reference():
    g = load_golden('readers')
    arr = reader.RAW(RAW_FILE, 24, 32, '16-bit Unsigned')
    assert arr.dtype == np.dtype('>u2') and arr.shape == (24, 32)
    assert np.array_equal(arr, g['raw_be'])
    little = reader.RAW(RAW_FILE, 24, 32, 'u2', littleEndian=True)
    assert np.array_equal(little, g['raw_be'].astype(np.uint16).byteswap())
    short = reader.RAW(RAW_FILE, 24, 40, 'u2')                 # file shorter than width*height: second extent re-derived
    assert short.shape == (24, 32)


def test_elbin_reader_and_layout_match_reference():
    g = load_golden('readers')
    arrs, labels = reader.elbin(ELBIN_FILE)
    assert arrs.dtype == np.uint16 and np.array_equal(arrs, g['elbin_frames'])
    assert [l['exposure time[s]'] for l in labels] == list(g['elbin_times'])
    assert [l['current[A]'] for l in labels] == list(g['elbin_current'])
    assert [l['voltage[V]'] for l in labels] == list(g['elbin_voltage'])
    lay = reader.elbin_layout(ELBIN_FILE)
    assert (lay['offset'], lay['gap'], lay['frames'], lay['shape']) == (32, 20, 3, (24, 40))
    px = 24 * 40 * 2
    for i in range(3):
        o = lay['offset'] + i * (px + lay['gap'])
        assert np.array_equal(lay['bytes'][o:o + px].view(np.uint16).reshape(24, 40), g['elbin_frames'][i])


@pytest.mark.gpu
def test_chain_consumes_file_bytes_as_stored():
    """big-endian RAW frames and an elbin stack go to the device as file bytes; K1 swaps / skips headers in its load.
    Result == the oracle chain on the arrays the reference's readers return."""
    import torch
    if not torch.cuda.is_available():
        pytest.skip('needs a CUDA device')
    from imgprocessor_b200 import engine, synth
    from imgprocessor_b200.camera import CameraCalibration
    from oracle import models
    g = load_golden('readers')
    # RAW: (24, 32) big-endian frame through the public API, no host byte swap
    H, W = 24, 32
    dark, flat = synth.dark_map(H, W), synth.flat_map(H, W)
    cal = CameraCalibration()
    cal.addDarkCurrent(dark)
    cal.addFlatField(flat)
    arr = reader.RAW(RAW_FILE, 24, 32, '16-bit Unsigned')
    import contextlib, io
    with contextlib.redirect_stdout(io.StringIO()):
        out = cal.correct(arr, threshold=0.1)
    want, _ = models.correct_chain_f32(g['raw_be'], dark, flat, 0.1, 3, None)
    assert np.array_equal(out, want.astype(np.float64))
    outb = cal.correct_batch(np.stack([arr, arr]), threshold=0.1)
    assert np.array_equal(outb[1], want)
    # a wide big-endian frame takes the TMA streaming kernel with the swap specialisation
    H2, W2 = 48, 256
    raw2 = synth.scene(H2, W2, 3, np.uint16)
    e = engine.get_engine(H2, W2)
    d2, f2 = synth.dark_map(H2, W2), synth.flat_map(H2, W2)
    e.set_dark(d2)
    e.set_flat(f2)
    be = torch.from_numpy(raw2.byteswap()).cuda()
    with e.ingest(big_endian=True):
        o2, _ = e.pointwise_median(be, 0.1, 3)
    w2, _ = models.median_threshold_model(models.pointwise_model(raw2, d2, f2, True), 0.1, 3)
    assert np.array_equal(o2.cpu().numpy(), w2)
    # elbin: whole file image on the device, frames 20 bytes apart, first pixels at byte 32
    lay = reader.elbin_layout(ELBIN_FILE)
    Hh, Ww = lay['shape']
    e3 = engine.get_engine(Hh, Ww)
    d3, f3 = synth.dark_map(Hh, Ww), synth.flat_map(Hh, Ww)
    e3.set_dark(d3)
    e3.set_flat(f3)
    e3.set_lens(None, None, None)
    buf = torch.from_numpy(lay['bytes']).cuda()
    o3 = e3.correct_file_bytes(buf, lay['offset'], lay['frames'], gap=lay['gap']).cpu().numpy()
    for i in range(lay['frames']):
        w3, _ = models.correct_chain_f32(g['elbin_frames'][i], d3, f3, 0.1, 3, None)
        assert np.array_equal(o3[i], w3)
